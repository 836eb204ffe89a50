"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the
device time (cold-cache, serialised: compare SHARES, not absolutes), libtok8s kernels marked.

  python tools/launch_summary.py gpurun_out/launches.csv [--skip N] > profiles/..._summary.txt
"""
import collections
import csv
import re
import sys

OURS = re.compile(r"(local_kernel|local_tma_kernel|one_shot_kernel|two_shot_kernel|nvls_kernel|"
                  r"nvls_inplace_kernel|two_shot_inplace_kernel|arrive_kernel|bcast_kernel|barrier_bench)")


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    n = 0
    for r in rows[1:]:
        n += 1
        if n <= skip:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui].strip(), 1e-3)
        m = OURS.search(r[ki])
        key = ("libtok8s::" + m.group(1)) if m else re.sub(r"\(.*", "", r[ki])[:90]
        tot[key] += v
        cnt[key] += 1
    total = sum(tot.values())
    print("launches: %d (first %d skipped)   total device time: %.1f us" % (sum(cnt.values()), skip, total))
    ours = sum(v for k, v in tot.items() if k.startswith("libtok8s::"))
    print("libtok8s kernels: %d launches, %.1f us = %.4f %% of the listed device time" %
          (sum(c for k, c in cnt.items() if k.startswith("libtok8s::")), ours, 100 * ours / max(total, 1e-9)))
    print("%8s %7s %10s  %s" % ("share%", "count", "avg_us", "kernel"))
    for k, v in tot.most_common(25):
        print("%8.3f %7d %10.2f  %s" % (100 * v / total, cnt[k], v / cnt[k], k))
    for k, v in tot.items():
        if k.startswith("libtok8s::") and k not in dict(tot.most_common(25)):
            print("%8.3f %7d %10.2f  %s" % (100 * v / total, cnt[k], v / cnt[k], k))


if __name__ == "__main__":
    main()
