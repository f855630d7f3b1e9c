"""SASS opcode census of lib/libb200icp.so per kernel -> profiles/r2_sass_opcodes.txt.
Shows that the hot loops are what DESIGN.md says they are (packed FP32x2 math, REDUX, float64
state, system-scope release/acquire for the peer exchange) and that no tensor-core / TMA opcode is
claimed that is not there."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX", "FMNMX3", "REDUX", "DADD", "DFMA", "DMUL", "DSETP", "MUFU", "F2F",
         "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "LDGSTS", "UBLKCP", "SYNCS", "ATOM", "RED", "MEMBAR", "HMMA", "UTCHMMA"]


def main():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "icp_slam-yolo_b200", "lib", "libb200icp.so")],
                   cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, counts = None, {}
        for ln in txt.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", ln)
            if m and cur:
                op = m.group(1)
                counts[cur][op] += 1
                if op in ("LD", "ST", "LDG", "STG") and m.group(2) and "SYS" in m.group(2):
                    counts[cur][op + ".SYS"] += 1
        for k, c in counts.items():
            name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name).replace("(KernelArgs)", "")
            tot = sum(c.values())
            watch = "  ".join(f"{w}={c[w]}" for w in WATCH if c.get(w))
            sysops = "  ".join(f"{w}={c[w]}" for w in c if w.endswith(".SYS"))
            out.append((f, tot, f"{name[:90]:90s} total={tot:5d}  {watch}  {sysops}"))
    with open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"), "w") as fh:
        fh.write("# SASS opcode census of icp_slam-yolo_b200/lib/libb200icp.so (sm_100a), per kernel: tools/sass_census.py\n")
        last = None
        for f, tot, line in sorted(out, key=lambda t: (t[0], -t[1])):
            if f != last:
                fh.write(f"\n## {f}\n")
                last = f
            fh.write(line.rstrip() + "\n")
    print(open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt")).read()[:3000])


if __name__ == "__main__":
    main()
