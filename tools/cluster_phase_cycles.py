"""Developer aid: where the cycles of the clustering kernels go (DUET_CL_DBG=1 makes the library print per-kernel serial
times and the per-phase cycle counts of k_cl_bucket to stderr).
    gpurun -- 'DUET_CL_DBG=1 python tools/cluster_phase_cycles.py'"""
import sys, numpy as np
sys.path.insert(0, '.')
from duet_b200 import synth
from duet_b200.sv_clustering import cluster_signatures
cols = synth.make_signatures(0, 2_000_000)
for _ in range(2):
    got, n, ms = cluster_signatures(*cols)
    print("ms", ms, "clusters", n, flush=True)
