"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void nfdpm::", "")[:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T:.1f} us of device time (cold-cache, serialised: compare shares)")
    print(f"{'share':>6} {'total_us':>10} {'n':>5} {'avg_us':>8}  kernel")
    for k, v in tot.most_common():
        print(f"{v / T * 100:5.1f}% {v:10.1f} {cnt[k]:5d} {v / cnt[k]:8.2f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
