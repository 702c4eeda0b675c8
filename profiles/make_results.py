#!/usr/bin/env python
"""profiles/r01_results.md from the perf-probe logs that came back in gpurun_out/ (scratch).

    python profiles/make_results.py <kernels.log> <with_row_reorder.log> <ref_gpu_baseline.log>

The logs are the stdout of tests/perf_probe.py (one JSON line per workload and K) and tests/ref_gpu_baseline.py."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load(fn):
    out = []
    for l in open(fn):
        try:
            out.append(json.loads(l))
        except ValueError:
            pass
    return out


def line(d, note=""):
    x = d["d0.3"]
    return "| %s | %d×%d | %d | %d | %.1f | %.0f | %.1f | %.1f | %.1f | %d/%d | %.2f / %.2f / %.2f | %s |" % (
        d["workload"], d["M"], d["N"], d["nnz"], d["K"], x["hot_ms"] * 1e3, x["hot_gflops"], x["cold_wide_ms"] * 1e3,
        x["hot_nowide_ms"] * 1e3, d["csr_order"]["hot_ms"] * 1e3, x["wide"][0], x["wide"][1], d["row_ms"], x["col_ms"], x["fmt_ms"], note)


def main():
    kernels, rows, ref = sys.argv[1:4]
    hdr = ("| workload | shape | nnz | K | default call, back to back µs | GFLOP/s | wide kernel alone, L2 flushed µs (2.6 = no wide groups) | "
           "BSMR split (NO_WIDE) µs | CSR-order kernel µs | wide/all row groups | row / col / format ms (warm) | note |\n" + "|---" * 12 + "|")
    out = ["# Round 1 measurements on B200 (sm_100a)", "",
           "All numbers from `gpurun` boxes (one B200 each unless noted), CUDA events on the launching stream, after warm-up.",
           "`tests/perf_probe.py <workloads> [--no-row]` prints one JSON line per (workload, K); `--no-row` keeps the original row order",
           "(`BSMR_ROW_IDENTITY`) so that the clustering time does not dominate the run.  \"default call\" = `bsmr_sddmm` with",
           "`BSMR_SDDMM_DEFAULT` (the per-K execution plan of DESIGN.md §3.0), 100 iterations back to back (L2-resident where the",
           "working set fits: what the reference's own timing loop measures, `src/sddmmKernel.cu:2565`).", "",
           "## Kernels, identity row order (`python tests/perf_probe.py nips mask70 mask90 mask98 blocks16k graph17 graph20 --no-row`)", "", hdr]
    out += [line(d) for d in load(kernels)]
    out += ["", "## With the BSMR row reorder (alpha = 0.3, reference_compat) (`python tests/perf_probe.py nips graph17`)", "", hdr]
    out += [line(d, "%d clusters" % d["clusters"]) for d in load(rows)]
    out += ["", "## Comparators on the same B200 (`python tests/ref_gpu_baseline.py nips 32 128`: the reference rebuilt for sm_100 in `oracle/_ref`, "
            "each call in its own process)", "", "| what | K | ms per SDDMM | GFLOP/s | note |", "|---|---|---|---|---|"]
    for d in load(ref):
        if d.get("mode") == "cusparse":
            out.append("| cuSPARSE `cusparseSDDMM` (the reference's comparator, `cuSparseBaseline/`) | %d | %.4f | %.0f | %d mismatches vs CPU |" % (
                d["K"], d["ms"], d["gflops"], d["mismatches"]))
        elif "ms" in d:
            out.append("| reference `sddmm_gpu` (BSMR kernels, wmma) | %d | %.4f | %.0f | row reorder %.0f ms, col reorder %.1f ms, %d clusters, %d mismatches vs CPU |" % (
                d["K"], d["ms"], d["gflops"], d["row_ms"], d["col_ms"], d["clusters"], d["mismatches"]))
    out += ["", "The reference's K > 32 kernels produce no output on sm_100 (`1 << tId` shuffle mask, `src/sddmmKernel.cu:2096`; DESIGN.md §4), "
            "hence 0 GFLOP/s and every value mismatching at K = 128.", ""]
    out += ["## Graphs far larger than L2 (r01h): see `profiles/r01h_l2_policy_sweep.md`", "",
            "R-MAT 2²³ rows, 2.5·10⁸ nnz, K=256 (BASELINE configs[4]): 27.95 ms = 4.58 TFLOP/s on one B200 (L2 policy on; 30.41 ms",
            "without), 15.38 / 7.64 ms on 2 / 4 GPUs (98.8 / 99.6 % of linear).  Residual kernel on R-MAT 2²², K=256 under ncu:",
            "5.49 TB/s of DRAM traffic = 84 % of the measured copy peak.", ""]
    open(os.path.join(HERE, "r01_results.md"), "w").write("\n".join(out))


if __name__ == "__main__":
    main()
