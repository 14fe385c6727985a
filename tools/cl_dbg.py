import sys, numpy as np
sys.path.insert(0, '.')
from duet_b200 import synth
from duet_b200.sv_clustering import cluster_signatures
cols = synth.make_signatures(0, 2_000_000)
for _ in range(2):
    got, n, ms = cluster_signatures(*cols)
    print("ms", ms, "clusters", n, flush=True)
