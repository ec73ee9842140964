#!/usr/bin/env python
"""Group a kernel's SASS by execution count (== by loop) and show instructions/stall samples per group.
    python profiles/ncu_loops.py gpurun_out/prof.ncu-rep [exec_count_to_dump ...]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
d = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    if len(r) >= 6 and r[0].startswith("0x"): d.append(r)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[5]) for r in d) or 1; tots = sum(int(r[2]) for r in d) or 1
g, n, s = collections.Counter(), collections.Counter(), collections.Counter()
for r in d:
    k = int(r[5]); g[k] += k; n[k] += 1; s[k] += int(r[2])
print("total instr", tot, "samples", tots)
for k, v in sorted(g.items(), key=lambda kv: -kv[1])[:18]:
    print("exec %8d: %4d instrs -> %9d (%4.1f%%)  samples %5d (%4.1f%%)" % (k, n[k], v, 100 * v / tot, s[k], 100 * s[k] / tots))
for want in sys.argv[2:]:
    print("---- exec == %s" % want)
    for r in d:
        if r[5] == want:
            st = sorted([(int(r[i] or 0), hdr[i][6:]) for i in stall_cols], reverse=True)[:2]
            print(r[2].rjust(4), r[1].strip()[:64].ljust(64), " ".join("%s:%d" % (a, b) for b, a in st if b > 0))
