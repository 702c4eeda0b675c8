#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum per launch, csv) -> text summary.

    python profiles/summarize_launches.py profiles/r01f_launch_list_bench_steps5.csv > profiles/r01f_launch_list_summary.txt"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ix = {h: j for j, h in enumerate(rows[start])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 2:]:
        if len(r) < len(ix):
            continue
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        n = r[ix["Kernel Name"]].split("(")[0][-70:]
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --quick --steps 5 --warmup 3   (B200, after the same command exited 0 without ncu)")
    print("# (the first 400 launches of the run: the four reorders, bsmr_plan_autotune's passes -- the explicit measurement bench.py makes before the clock starts -- and the warm-up / timed steps)")
    print("# full list: %s.  Times are ns, cold-cache and serialised: compare shares." % sys.argv[1])
    print("# The timed step of bench.py launches ONE kernel on the nips workload (wide_sddmm_kernel: all six row groups are wide), so its share of the step is 100 %.")
    print("# Everything else below is the reorder (4 calls: first + 3 warm; bsa_cluster_kernel is the clustering chain), the L2 flush (FillFunctor),")
    print("# the autotune passes (wide / dense-block / residual kernels, four passes per candidate) and the per-kernel / hot-loop passes of the bench.")
    print("%-72s %6s %14s %7s" % ("kernel", "count", "total ns", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %6d %14.0f %6.2f%%" % (n, c, t, 100 * t / tot))
    print("%-72s %6d %14.0f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


if __name__ == "__main__":
    main()
