"""Top SASS instructions by warp-stall samples for one kernel of an ncu report (needs --import-source on / -lineinfo).

    python tools/ncu_hot.py gpurun_out/r02e/prof.ncu-rep lbs_vertex_bwd [N]
"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
head = rows[hi]
isrc, isamp, iexec = head.index("Source"), head.index("# Samples"), head.index("Instructions Executed")
stalls = [(i, h) for i, h in enumerate(head) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[hi + 1:]:
    if len(r) != len(head) or r[0] == "Address":
        break                                   # next kernel instance
    body.append(r)
tot = sum(int(r[isamp] or 0) for r in body)
print("%s: %d SASS instructions, %d samples" % (pat, len(body), tot))
agg = {}
for i, h in stalls:
    agg[h] = sum(int(r[i] or 0) for r in body)
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in sorted(enumerate(body), key=lambda t: -int(t[1][isamp] or 0))[:top]:
    s = int(r[isamp] or 0)
    why = sorted(((int(r[i] or 0), h[6:]) for i, h in stalls), reverse=True)[:2]
    print("%5d  %5.1f%%  x%-8s %-70s %s" % (n, 100.0 * s / max(tot, 1), r[iexec], r[isrc][:70], " ".join("%s=%d" % (w, c) for c, w in why if c)))
