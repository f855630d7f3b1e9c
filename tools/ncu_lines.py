"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report.

ncu's CSV export of the source page is SASS only; this joins it with `nvdisasm -g` of the same
kernel (same instruction order) to attribute executed instructions and samples to CUDA lines.

  python tools/ncu_lines.py gpurun_out/x.ncu-rep icp_align_pair_kernelILi2ELb1ELi2 [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ii, si, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    data = []
    for r in rows[h + 1:]:
        if len(r) <= max(ii, si):
            continue
        f = lambda x: int(x.replace(",", "")) if x else 0
        data.append((r[src].strip(), f(r[ii]), f(r[si])))
    return data


def line_map(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("B200ICP_LIB", os.path.join(ROOT, "icp_slam-yolo_b200", "lib", "libb200icp.so"))],
                   cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel_substr not in txt:
            continue
        lines, cur, on, out = txt.splitlines(), None, False, []
        for ln in lines:
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                on = kernel_substr in m.group(1)
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                out.append(cur)
        return out
    raise SystemExit("kernel not found in the library")


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    sass, lm = sass_rows(rep), line_map(kern)
    if len(sass) != len(lm):
        print(f"warning: {len(sass)} SASS rows in the report, {len(lm)} in the library (rebuilt since?)")
    agg = {}
    for (txt, n, s), loc in zip(sass, lm):
        a = agg.setdefault(loc, [0, 0])
        a[0] += n; a[1] += s
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    src = {}
    print(f"total warp-instructions {ti}, samples {ts}")
    for loc, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            path = os.path.join(ROOT, "icp_slam-yolo_b200", "csrc", loc[0])
            if path not in src and os.path.exists(path):
                src[path] = open(path).read().splitlines()
            if path in src and loc[1] <= len(src[path]):
                text = src[path][loc[1] - 1].strip()[:110]
        print(f"{(loc[0] + ':' + str(loc[1])) if loc else '?':>22} {100 * n / ti:6.2f}% instr {100 * s / max(ts, 1):6.2f}% samples  {text}")


if __name__ == "__main__":
    main()
