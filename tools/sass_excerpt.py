"""Per-kernel SASS evidence for profiles/: opcode histogram (cuobjdump -sass of the built library) and the lines that
carry the mnemonics a reader looks for (tcgen05 = UTC*, TMA = UTMA*, mbarrier = SYNCS, MUFU, packed fp32 = *2, REDUX, ...).

    python tools/sass_excerpt.py <kernel-substring> [mnemonic-regex] > profiles/rNN_sass_<kernel>.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200", "ltr_b200", "lib", "libltr_b200.so")


def main():
    sub = sys.argv[1]
    want = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"UTC|UTMA|SYNCS|LDTM|STTM|ELECT|MUFU|REDUX|UBLKPF|BAR\.|UCGABAR|NANOSLEEP")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, ops, lines, regs = None, collections.Counter(), [], ""
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if fn and sub in fn:
                break
            fn = m.group(1)
            ops, lines = collections.Counter(), []
            continue
        if fn is None or sub not in fn:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            body = m.group(2).strip()
            op = re.sub(r"^@!?U?P\d+\s+", "", body).split()[0]
            ops[op] += 1
            if want.search(op):
                lines.append(f"/*{m.group(1)}*/ {body}")
    demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(f"# {demangled}\n# {sum(ops.values())} SASS instructions ({sum(ops.values()) * 16 / 1024:.1f} KB), sm_100a, from {os.path.relpath(LIB, ROOT)}")
    fam = collections.Counter()
    for op, c in ops.items():
        fam[op.split(".")[0]] += c
    print("# opcode families: " + ", ".join(f"{k} {v}" for k, v in fam.most_common(40)))
    print("# full opcodes matching the filter: " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items()) if want.search(k)))
    print(f"# first 60 of {len(lines)} matching instructions:")
    for ln in lines[:60]:
        print(ln)


if __name__ == "__main__":
    main()
