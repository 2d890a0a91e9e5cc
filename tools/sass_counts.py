"""Per-kernel SASS mnemonic counts of libdiff3d_b200.so (the Blackwell-native instruction evidence):
tcgen05 MMAs (UTCHMMA = kind::f16, UTCQMMA = kind::f8f6f4, UTCOMMA = kind::mxf4 block-scaled), TMEM traffic (LDTM / STTM /
UTCCP), TMA (UTMALDG / UTMASTG), legacy tensor path (HMMA) and packed fp32 (FFMA2 / FMUL2 / FADD2).

    python tools/sass_counts.py > profiles/<tag>_sass_counts.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diff3dhpe_b200", "libdiff3d_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "FFMA2", "FMUL2",
        "FADD2", "F2FP", "MUFU", "SHFL"]


def main():
    exe = "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    rows = []
    for chunk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = chunk.split("\n", 1)[0].strip()
        ops = collections.Counter()
        total = 0
        for line in chunk.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                total += 1
                ops[m.group(1)] += 1
        rows.append((name, total, ops))
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
    print("# SASS mnemonic counts per kernel (`cuobjdump -sass diff3dhpe_b200/libdiff3d_b200.so`, sm_100a)\n")
    print("| kernel | instrs | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    tot = collections.Counter()
    for (mangled, total, ops), nice in zip(rows, names):
        nice = re.sub(r"\(anonymous namespace\)::", "", nice)
        nice = re.sub(r"^void d3d::", "", nice)
        nice = re.sub(r"\(.*$", "", nice)
        if not any(ops[k] for k in KEYS[:10]) and "kernel" not in nice:
            continue
        print(f"| `{nice}` | {total} | " + " | ".join(str(ops[k]) if ops[k] else "" for k in KEYS) + " |")
        for k in KEYS:
            tot[k] += ops[k]
    print("| **total** | | " + " | ".join(str(tot[k]) for k in KEYS) + " |")


if __name__ == "__main__":
    main()
