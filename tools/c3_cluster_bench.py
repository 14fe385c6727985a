import sys
sys.path.insert(0, ".")
from duet_b200 import synth
from duet_b200.sv_clustering import cluster_signatures
cols = synth.make_signatures(0, n=2_000_000)
for i in range(3):
    ids, nc, ms = cluster_signatures(*cols)
    print("C3 2M signatures:", nc, "clusters, device ms", ms)
