"""SASS opcode summary of the hot kernels in the built library (no GPU needed):
    python tools/sass_summary.py [kernel-substring ...] > profiles/rNN_sass_summary.txt
Per kernel: instruction count, registers, and the opcodes that prove what the kernel does -- UBLKCP (cp.async.bulk: the TMA
bulk copies), SYNCS (mbarrier), LDS/STS (shared-memory image painting), LDG/STG (global), ATOMS/RED, PRMT, POPC, SHFL, BAR."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "snakes_b200", "libsnk.so")
want = sys.argv[1:] or ["k_step_lane<2, 0, 2>", "k_lane_logic<2, 0>", "k_lane_paint2<2, 0, 2>", "k_lane_logic<3, 2>", "k_lane_paint2<3, 2, 3>",
                        "k_step_rows<2>", "k_upscale84<6, 4>", "k_extract_main<6>", "k_scripted_actions<2>", "k_gae"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
regs = {}
for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur][m.group(1).split(".")[0]] += 1
        funcs[cur]["_total"] += 1
keys = ["UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "PRMT", "POPC", "SHFL", "BAR", "WARPSYNC", "IMAD", "LOP3", "BRA", "CALL"]
print("%-34s %6s %4s %5s  %s" % ("kernel (sm_100a SASS)", "instrs", "regs", "stack", "  ".join("%s" % k for k in keys)))
for name, c in funcs.items():
    d = demangle(name)
    short = re.sub(r"^void |\(.*$", "", d)
    if not any(w in short for w in want):
        continue
    r = regs.get(name, (0, 0, 0))
    print("%-34s %6d %4d %5d  %s" % (short, c["_total"], r[0], r[1], "  ".join("%*d" % (len(k), c[k]) for k in keys)))
