#!/usr/bin/env python
"""Hottest SASS instructions of one kernel in an ncu report: python tools/ncu_hot.py <rep> <kernel> [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, h0 in enumerate(hi[:1]):
    h = rows[h0]; ci = {n: i for i, n in enumerate(h)}
    body = [r for r in rows[h0 + 1:(hi[k + 1] - 1 if k + 1 < len(hi) else None)] if len(r) == len(h)]
    tot = sum(int(r[ci["# Samples"]]) for r in body)
    inst = sum(int(r[ci["Instructions Executed"]]) for r in body)
    print(f"samples {tot}  warp instructions {inst}")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    for n, r in sorted(enumerate(body), key=lambda x: -int(x[1][ci["# Samples"]]))[:top]:
        s = int(r[ci["# Samples"]])
        why = sorted(((int(r[ci[c]]), c[6:]) for c in stalls), reverse=True)[:2]
        print(f"{n:4d} {100*s/max(tot,1):5.1f}%  x{r[ci['Instructions Executed']]:>8}  {r[ci['Source']].strip()[:70]:70s} {why}")
