"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md), from the
built library:  python tools/sass_counts.py > profiles/rNN_sass_counts.md"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "flair_b200" / "libflair_b200.so"
COLS = OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"),
                    ("SYNCS", r"\bSYNCS"), ("ELECT", r"\bELECT"), ("STG.256", r"\bSTG\.E\.(ENL2\.)?256"),
                    ("LDG.256", r"\bLDG\.E\.(ENL2\.)?256"), ("HMMA", r"\bHMMA")])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    total = len(re.findall(r"UTCHMMA|UTMALDG|LDTM", sass))
    kernels, cur = OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {c: 0 for c in COLS}
            kernels[cur]["n"] = 0
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            kernels[cur]["n"] += 1
            for c, pat in COLS.items():
                if re.search(pat, line):
                    kernels[cur][c] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence of the tcgen05 / TMEM / TMA path (`cuobjdump -sass flair_b200/libflair_b200.so`, sm_100a, "
          "`python tools/sass_counts.py`)\n")
    print(f"`cuobjdump -sass flair_b200/libflair_b200.so | grep -c 'UTCHMMA\\|UTMALDG\\|LDTM'` -> **{total}**\n")
    print("UTCHMMA = tcgen05.mma (kind::f16), UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld (TMEM -> registers), "
          "UTCBAR = tcgen05.commit,\nSYNCS = mbarrier ops, ELECT = elect.sync, STG/LDG.256 = 32-byte global store / load, "
          "HMMA = legacy mma.sync (none).  `instr` = SASS instructions of the kernel.\n")
    print("| kernel | instr | " + " | ".join(COLS) + " |\n|---|---:|" + "---:|" * len(COLS))
    rows = []
    for (k, d), n in zip(kernels.items(), names):
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", re.sub(r"\(.*$", "", n))
        rows.append((n, d))
    tc = [r for r in rows if r[1]["UTCHMMA"] or r[1]["UTMALDG"] or r[1]["LDTM"]]
    for n, d in sorted(tc):
        print(f"| `{n}` | {d['n']} | " + " | ".join(str(d[c]) for c in COLS) + " |")
    rest = [r for r in rows if r not in tc]
    hm = sum(d["HMMA"] for _, d in rows)
    print(f"\n{len(rest)} further kernels (HBM- / latency-bound, SIMT) carry none of these; HMMA in the whole library: {hm}.")


if __name__ == "__main__":
    sys.exit(main())
