"""Aggregate an ncu source-page CSV into source-line ranges delimited by '// ----' marker comments.

    python tools/ncu_regions.py src.csv <lib.so> <kernel-substring> <source-file-basename>
"""
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as N  # noqa: E402


def main():
    src_csv, so_path, ksub, fname = sys.argv[1:5]
    table = N.line_table(so_path, ksub)
    rows = list(csv.reader(open(src_csv)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    b = blocks[0]
    hdr = b["hdr"]
    ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
    fn = [k for k in table if len(table[k]) == len(b["rows"])][0]
    here = os.path.dirname(os.path.abspath(so_path))
    path = os.path.join(here, "..", "..", "csrc", fname)
    lines = open(path).read().splitlines()
    marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(lines)
             if l.strip().startswith("// ----") or l.strip().startswith("template <") or "__device__" in l or "__global__" in l]
    def region(ln):
        name = "(top)"
        for m, text in marks:
            if m <= ln:
                name = f"{m}: {text}"
            else:
                break
        return name
    agg, smp = collections.Counter(), collections.Counter()
    tot = ts = 0
    for (addr, line, sass), r in zip(table[fn], b["rows"]):
        i, s = int(r[ie]), int(r[ns])
        tot += i
        ts += s
        key = region(line[1]) if line and line[0] == fname else (line[0] if line else "?")
        agg[key] += i
        smp[key] += s
    for k, v in agg.most_common():
        print(f"{v:9d} {v / tot:6.3f} | smp {smp[k]:5d} {smp[k] / max(1, ts):6.3f} | {k}")
    print(tot, ts)


if __name__ == "__main__":
    main()
