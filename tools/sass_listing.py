"""Trimmed SASS listing of one kernel of the built library (addresses + instructions, encodings stripped), with the
mnemonic counts that prove which units it drives (UTCHMMA / UTCQMMA = tcgen05.mma, UBLKCP = cp.async.bulk, SYNCS = mbarrier).

    python tools/sass_listing.py lbs_blend_fwd_bf3 > profiles/<round>_sass_blend_fwd_bf3.txt
"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from psi_release_b200 import build  # noqa: E402

pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", build.lib_path()], capture_output=True, text=True).stdout
out, name, keep = [], None, False
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        keep = pat in m.group(1) and name is None
        if keep:
            name = m.group(1)
        continue
    if keep:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((m.group(1), m.group(2).strip()))
if not out:
    sys.exit("no kernel matching %r" % pat)
cnt = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0] for _, ins in out)
print("# cuobjdump -sass, %s (addresses + instructions; encodings stripped)" % name)
print("# %d instructions; mnemonic counts: %s" % (len(out), ", ".join("%s %d" % kv for kv in cnt.most_common(14))))
tc = {k: v for k, v in cnt.items() if k in ("UTCHMMA", "UTCQMMA", "UTCMMA", "UBLKCP", "SYNCS", "UTCBAR", "LDTM", "UTMALDG", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "HMMA", "LDSM")}
print("# tensor / TMA / packed-math mnemonics: %s" % (", ".join("%s %d" % kv for kv in sorted(tc.items())) or "none"))
for a, ins in out:
    print("%s  %s" % (a[-4:], ins))
