"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [top]

ncu's `--page source --csv` lists SASS instructions (in address order) with 'Instructions
Executed' and stall '# Samples'; `nvdisasm --print-line-info` on the cubin gives the source line of
every SASS instruction of the same function in the same order.  The two are zipped.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel_sub):
    # the library is linked from several objects whose cubins share one name: disassemble per object
    bdir = os.path.join(ROOT, "snakes_b200", "csrc", "_build")
    objs = sorted(os.path.join(bdir, f) for f in os.listdir(bdir) if f.endswith(".o")) if os.path.isdir(bdir) else []
    out = ""
    for obj in objs or [os.path.join(ROOT, "snakes_b200", "libsnk.so")]:
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for f in sorted(os.listdir(tmp)):
            if f.endswith(".cubin"):
                out += subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    funcs, cur, line = {}, None, None
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            line = None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            line = (os.path.basename(m.group(1)), int(m.group(2)), (os.path.basename(inl.group(1)), int(inl.group(2))) if inl else None)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.+?);", l)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip(), line))
    for name, ins in funcs.items():
        if kernel_sub in name:
            return name, ins
    raise SystemExit("kernel not found: %s (have %s)" % (kernel_sub, list(funcs)))


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    name, ins = sass_lines(ksub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # the report may hold several kernels: take the section whose SASS has as many instructions as the cubin's function
    hdr, sections = None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            sections.append([])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and sections and len(r) == len(hdr):
            sections[-1].append(r)
    inst = next((sec for sec in sections if len(sec) == len(ins)), sections[0] if sections else [])
    iI, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    assert len(inst) == len(ins), (len(inst), len(ins))
    by_line, by_line_s = {}, {}
    tot = sum(int(r[iI]) for r in inst)
    tots = sum(int(r[iS]) for r in inst)
    for r, (addr, text, line) in zip(inst, ins):
        key = line[:2] if line else ("?", 0)
        by_line[key] = by_line.get(key, 0) + int(r[iI])
        by_line_s[key] = by_line_s.get(key, 0) + int(r[iS])
    print("kernel %s: %d SASS instrs, %d warp-instructions executed, %d stall samples" % (name, len(ins), tot, tots))
    src_cache = {}

    def src(f, n):
        if f not in src_cache:
            p = os.path.join(ROOT, "snakes_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        return src_cache[f][n - 1].strip()[:90] if 0 < n <= len(src_cache[f]) else ""

    print("---- top lines by instructions executed")
    for key, v in sorted(by_line.items(), key=lambda kv: -kv[1])[:top]:
        print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * v / tot, 100.0 * by_line_s[key] / max(tots, 1), key[0], key[1], src(*key)))
    print("---- top lines by stall samples")
    for key, v in sorted(by_line_s.items(), key=lambda kv: -kv[1])[:top // 2]:
        print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100.0 * v / max(tots, 1), 100.0 * by_line[key] / tot, key[0], key[1], src(*key)))


if __name__ == "__main__":
    main()
