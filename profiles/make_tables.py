#!/usr/bin/env python
"""Regenerates the measured tables of BASELINE.md (section 4) and DESIGN.md (section 5) from the committed bench lines.

    python profiles/make_tables.py profiles/r02o_bench_default.json

Prints markdown; the numbers in the two documents are pasted from this output."""
import json
import sys


def us(ms):
    return "%.1f" % (ms * 1e3) if ms < 0.1 else "%.0f" % (ms * 1e3)


def main():
    d = json.load(open(sys.argv[1]))
    comps = {}
    for c in d.get("comparators") or []:
        comps[(c["workload"], c["K"], c["comparator"])] = c
    rows = [("nips K=128 (configs[1], headline)", "nips", 128, d["ms_per_step"], d["kernels"]["step_ms_hot_l2"], d["value"], d["roofline"]["step"]["frac"],
             d["execution_plan"], d["reorder"]["row_ms"], d["reorder"]["col_ms"], d["reorder"]["format_ms"], d["reorder"]["clusters"], None)]
    for c in d["configs"]:
        wl = "nips" if c["name"].startswith("nips") else ("mask%s" % c["name"].split()[1] if c["name"].startswith("mask") else None)
        rows.append((c["name"], wl, c["K"], c["ms_per_step"], c["ms_hot_l2"], c["value"], c["roofline"]["frac"], c["execution_plan"],
                     c["reorder"]["row_ms"], c["reorder"]["col_ms"], c["reorder"]["format_ms"], c["reorder"]["clusters"], c.get("fp16_b")))
    print("| config | cold us (GFLOP/s, % of HBM roofline) | hot us (GFLOP/s) | plan | reorder: row + col + format ms (clusters) | cuSPARSE hot us | reference sddmm_gpu hot us |")
    print("|---|---|---|---|---|---|---|")
    for name, wl, K, cold, hot, gf, frac, plan, rms, cms, fms, cl, f16 in rows:
        nnzK = gf * cold  # = 2 nnz K / 1e6
        cs = comps.get((wl, K, "cusparse"))
        rf = comps.get((wl, K, "bsmr_ref"))
        extra = "; fp16 B: %s (%.0f)" % (us(f16["ms_per_step"]), f16["value"]) if f16 else ""
        print("| %s | %s (%.0f, %.1f %%)%s | %s (%.0f) | %s | %.1f + %.1f + %.1f (%d) | %s | %s |" % (
            name, us(cold), gf, 100 * frac, extra, us(hot), nnzK / hot, plan.split(" (")[0], rms, cms, fms, cl,
            "%.1f (%.0f)" % (cs["ms"] * 1e3, cs["gflops"]) if cs else "-",
            ("%.1f (%.0f)" % (rf["ms"] * 1e3, rf["gflops"]) if rf["mismatches_vs_sddmm_cpu"] == 0 else "writes nothing") if rf else "-"))
    print()
    for b in d.get("batch") or []:
        print("batch: %s: loop %.0f us, batched %.0f us, %.2fx, %.0f GFLOP/s" % (b["workload"], b["loop_of_single_calls_ms"] * 1e3, b["batched_ms"] * 1e3, b["speedup"], b["gflops_batched"]))
    print("e2e nips: %.0f us pipelined (%.0f GFLOP/s), %.0f us blocking; cpu_baseline %.2f GFLOP/s on %d threads" % (
        d["e2e"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["blocking_call_ms"] * 1e3, d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))


if __name__ == "__main__":
    main()
