#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel.

    python scripts/sass_summary.py [path/to/libeasyrag_b200.so] > profiles/sass_summary.txt

Runs ``cuobjdump -sass`` on the built library (no GPU needed) and prints, for every kernel, how many
UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG / UBLKCP (TMA), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
HMMA (legacy mma.sync), LDGSTS (cp.async), ATOMS / ATOMG / RED instructions it contains, plus its code size.
Mapping PTX -> SASS: /opt/skills/guides/B200_PROFILING.md "What proves a Blackwell-native kernel".
"""
from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COLS = ["UTC*MMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "ATOMS", "ATOMG", "RED", "BAR", "SYNCS"]


def classify(op: str):
    if re.match(r"UTC[A-Z]*MMA", op):
        return "UTC*MMA"
    for c in COLS[1:]:
        if op.startswith(c):
            return c
    return None


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
        return out.splitlines()
    except Exception:
        return names


def main():
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "easyrag_b200" / "_lib" / "libeasyrag_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    kernels: "OrderedDict[str, Counter]" = OrderedDict()
    sizes = {}
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            sizes[cur] = 0
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            sizes[cur] = max(sizes[cur], int(m.group(1), 16) + 16)
            c = classify(m.group(2).split(".")[0] if not m.group(2).startswith("UTC") else m.group(2).split(".")[0])
            if c:
                kernels[cur][c] += 1
    names = demangle(list(kernels))
    print(f"# SASS summary of {lib.relative_to(ROOT) if lib.is_relative_to(ROOT) else lib} (cuobjdump -sass; counts of instructions per kernel)")
    print("# UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG/UBLKCP = TMA, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, HMMA = mma.sync")
    print("kernel | bytes | " + " | ".join(COLS))
    for (mangled, cnt), name in zip(kernels.items(), names):
        short = re.sub(r"\(.*$", "", name)
        short = short.replace("ezr::", "").replace("void ", "")
        print(f"{short} | {sizes[mangled]} | " + " | ".join(str(cnt.get(c, 0)) for c in COLS))


if __name__ == "__main__":
    main()
