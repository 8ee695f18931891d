"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

  python tools/summarize_launches.py gpurun_out/launches.csv [--last N] [--skip N] > profiles/xxx.md

--last N keeps only the last N launches (e.g. one forward), --skip N drops the first N (setup).
ncu per-launch times are cold-cache and serialised: read the SHARES, not the absolute sum.
"""
import collections
import csv
import sys


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        rows.append((r[ik], us))
    return rows


def short(name):
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    cut = name.find("(")
    return (name if cut < 0 else name[:cut])[:90]


def main():
    args = sys.argv[1:]
    path = args[0]
    rows = load(path)
    if "--skip" in args:
        rows = rows[int(args[args.index("--skip") + 1]):]
    if "--last" in args:
        rows = rows[-int(args[args.index("--last") + 1]):]
    tot, cnt = collections.Counter(), collections.Counter()
    for k, us in rows:
        tot[short(k)] += us
        cnt[short(k)] += 1
    total = sum(tot.values())
    print(f"launches: {len(rows)}; sum of kernel durations: {total / 1e3:.1f} ms\n")
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in tot.most_common():
        print(f"| `{k}` | {cnt[k]} | {v:.0f} | {v / cnt[k]:.1f} | {100 * v / total:.1f}% |")


if __name__ == "__main__":
    main()
