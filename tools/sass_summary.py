#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS mnemonics in libmoshi_b200.so (cuobjdump -sass): tcgen05 MMAs (UTC*MMA),
TMEM loads (LDTM), TMA tensor / bulk copies (UTMALDG / UBLKCP), tcgen05 commits (UTCBAR) and TMEM allocation (UTCATOM*).
Writes profiles/sass_summary.md; runs on the build container (no GPU needed).

    python tools/sass_summary.py [path/to/lib.so]
"""
from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "moshi_b200" / "_C" / "libmoshi_b200.so"
PATTERNS = {"UTC*MMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA\b", "LDTM (tcgen05.ld)": r"\bLDTM\b", "UTMALDG (TMA tensor load)": r"\bUTMALDG\b",
            "UBLKCP (bulk copy)": r"\bUBLKCP\b", "UTCBAR (tcgen05.commit)": r"\bUTCBAR\b", "HMMA/IMMA (legacy mma.sync)": r"\b[HI]MMA\b",
            "FFMA": r"\bFFMA\b"}


def demangle(names: list[str]) -> dict[str, str]:
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts: dict[str, Counter] = defaultdict(Counter)
    kinds: dict[str, Counter] = defaultdict(Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        for label, pat in PATTERNS.items():
            if re.search(pat, line):
                counts[cur][label] += 1
        mk = re.search(r"\b(UTC[A-Z]*MMA)\b", line)
        if mk:
            kinds[cur][mk.group(1)] += 1
    names = demangle(list(counts))
    rows = []
    for k, c in counts.items():
        short = names[k].replace("(anonymous namespace)::", "").replace("b200::", "")
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", short))
        rows.append((short, c, kinds[k]))
    # one row per distinct kernel name (static kernels are duplicated across translation units)
    seen, uniq = set(), []
    for r in sorted(rows, key=lambda r: (-r[1]["UTC*MMA (tcgen05.mma)"], -r[1]["UTMALDG (TMA tensor load)"], r[0])):
        if r[0] in seen:
            continue
        seen.add(r[0])
        uniq.append(r)
    labels = list(PATTERNS)
    lines = ["# SASS evidence: Blackwell-only instructions per kernel", "",
             f"`cuobjdump -sass {LIB.relative_to(ROOT)}` (sm_100a), counted by `tools/sass_summary.py`.  `UTC*MMA` = `tcgen05.mma` "
             "(UTCHMMA: kind::f16 / tf32, UTCIMMA: kind::i8), `LDTM` = `tcgen05.ld`, `UTMALDG` = `cp.async.bulk.tensor` (TMA), `UBLKCP` = "
             "`cp.async.bulk`, `UTCBAR` = `tcgen05.commit`.  Kernels without any of them are SIMT (`FFMA` shown for scale).", "",
             "| kernel | " + " | ".join(labels) + " | MMA kinds |", "|---|" + "---:|" * len(labels) + "---|"]
    for short, c, kd in uniq:
        if not any(c[l] for l in labels[:6]) and c["FFMA"] < 50:
            continue
        lines.append(f"| `{short}` | " + " | ".join(str(c[l]) for l in labels) + " | " + ", ".join(f"{k} x{v}" for k, v in kd.items()) + " |")
    (ROOT / "profiles" / "sass_summary.md").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
