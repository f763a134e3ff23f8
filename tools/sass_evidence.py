"""Developer tool: SASS mnemonic counts per kernel of libb200gs.so (cuobjdump -sass), written to profiles/.
    python tools/sass_evidence.py profiles/r02b_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "sdp-gs_b200", "b200gs", "libb200gs.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
WATCH = ["UBLKCP", "LDGSTS", "REDG", "RED", "ACQBULK", "PREEXIT", "MATCH", "REDUX", "MUFU.EX2", "MUFU.RCP", "FFMA2", "FMUL2", "FADD2", "ATOM", "ATOMG",
         "BAR.SYNC", "SHFL", "VOTE", "MEMBAR", "ERRBAR", "CCTL", "SYNCS", "HMMA", "UTCMMA", "UTCHMMA"]
out = ["SASS mnemonic counts per kernel of sdp-gs_b200/b200gs/libb200gs.so (cuobjdump -sass, sm_100a), end of round 2.",
       "UBLKCP = TMA bulk copy (cp.async.bulk), LDGSTS = cp.async, REDG/RED = red.global.add (v4.f32 in the blend backward),",
       "ACQBULK/PREEXIT = programmatic dependent launch (griddepcontrol.wait / launch_dependents), MATCH = match.any (radix ranking),",
       "FFMA2/FMUL2/FADD2 = packed f32x2 arithmetic (fma/mul/add.rn.f32x2: the blend kernels' pair evaluation and channel sums),",
       "MUFU.EX2 = exp2 of expf, no HMMA / UTC*MMA anywhere (the path has no dense contraction).", ""]
name, counts, n = None, None, 0
def flush():
    if name:
        short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d*", "", name)
        out.append("%-70s %5d instr  %s" % (name[-70:], n, " ".join("%s=%d" % (k, counts[k]) for k in WATCH if counts.get(k))))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, n = subprocess.run(["c++filt", "-p", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip() or m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        n += 1
        op = m.group(1)
        for k in WATCH:
            if op == k or op.startswith(k + "."):
                counts[k] += 1
flush()
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out[:6] + [l for l in out[6:] if "blend" in l]))
