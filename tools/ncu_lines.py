"""Join an ncu source-page CSV (per-SASS-instruction counters) with nvdisasm line info, and print
executed warp-instructions and stall samples aggregated per CUDA source line.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv <lib.so> <kernel-substring> [top]
"""
import csv
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile


def line_table(so_path, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur_fn, cur_line, table = None, None, {}
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and kernel_sub in cur_fn:
                table.setdefault(cur_fn, []).append((int(m.group(1), 16), cur_line, m.group(2).strip()))
        if table:
            return table
    return {}


def main():
    src_csv, so_path, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(so_path, ksub)
    rows = list(csv.reader(open(src_csv)))
    # the CSV may hold several kernels: blocks start with a "Kernel Name" row
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    for blk in blocks[:1]:
        hdr = blk["hdr"]
        ie = hdr.index("Instructions Executed")
        ns = hdr.index("# Samples")
        fn = None
        for k in table:
            if len(table[k]) == len(blk["rows"]):
                fn = k
        if fn is None:
            print("could not match instruction count:", len(blk["rows"]), {k: len(v) for k, v in table.items()})
            return
        per_line = collections.defaultdict(lambda: [0, 0])
        tot_i = tot_s = 0
        for (addr, line, sass), r in zip(table[fn], blk["rows"]):
            i, s = int(r[ie]), int(r[ns])
            per_line[line][0] += i
            per_line[line][1] += s
            tot_i += i
            tot_s += s
        print(f"kernel {blk['name']}: {tot_i} warp-instructions, {tot_s} samples")
        srcs = {}
        for (f, l), (i, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
            if f not in srcs:
                cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(so_path)), "..", "..", "csrc", f))
                srcs[f] = open(cand[0]).read().splitlines() if cand else []
            text = srcs[f][l - 1].strip()[:90] if srcs[f] and l <= len(srcs[f]) else ""
            print(f"{i:9d} {i / tot_i:6.3f} | smp {s:6d} {s / max(1, tot_s):6.3f} | {f}:{l:<4d} {text}")


if __name__ == "__main__":
    main()
