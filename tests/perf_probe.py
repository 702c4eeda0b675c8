"""B200 box: timing overview of the hot path over several workloads (not a pytest test).

    python tests/perf_probe.py [nips] [mask90] [blocks] [graph17] [graph20] ...

For every workload and K it prints the reorder times, the split, and the SDDMM time hot (100 back-to-back
iterations, L2 resident when the working set fits) and cold (L2 flushed before every pass, the two kernels
timed separately), for alpha = delta = 0.3, delta = 0 (everything through tcgen05) and the CSR-order
residual kernel alone.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def workloads(pkg, names):
    s = pkg.synth
    table = {
        "nips": (lambda: s.nips_like(), [32, 128, 256]),
        "mask70": (lambda: s.dlmc_mask(0.70), [64]),
        "mask90": (lambda: s.dlmc_mask(0.90), [64]),
        "mask98": (lambda: s.dlmc_mask(0.98), [64]),
        "blocks": (lambda: s.block_structured(1000, 2000, seed=11, groups=12, cols_per_group=64), [128]),
        "blocks16k": (lambda: s.block_structured(16000, 16000, seed=5, groups=200, cols_per_group=96, noise=0.001), [128]),
        "graph17": (lambda: s.rmat(17, 3_000_000, 17), [128]),
        "graph20": (lambda: s.rmat(20, 30_000_000, 20), [128]),
    }
    for n in names:
        yield n, table[n][0](), table[n][1]


def main():
    import torch
    pkg = entry.load_package()
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["nips", "mask90", "blocks"]
    skip_row = "--no-row" in sys.argv
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for name, (M, N, ro, ci), Ks in workloads(pkg, names):
        nnz = len(ci)
        plan = pkg.Plan(ctx, M, N, ro, ci)
        big = M > 200000
        for rep in range(2):   # second call: scratch arena warm
            t0 = time.perf_counter()
            plan.row_reorder(0.3, flags=pkg.ROW_IDENTITY if (skip_row or big) else pkg.ROW_REFERENCE_COMPAT)
            row_wall = (time.perf_counter() - t0) * 1e3
            info_r = plan.info()
        for K in Ks:
            A, B = pkg.synth.make_ab(M, N, K)
            dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
            dP = torch.zeros(nnz, device="cuda")
            out = {"workload": name, "M": M, "N": N, "nnz": nnz, "K": K, "row_ms": info_r["row_reordering_ms"], "row_wall_ms": row_wall,
                   "clusters": info_r["num_clusters_true"], "block_size": info_r["block_size"], "cluster_ms": info_r["cluster_kernel_ms"]}
            for delta in (0.3, 0.0):
                for rep in range(2):
                    t0 = time.perf_counter()
                    plan.col_reorder(delta)
                    col_wall = (time.perf_counter() - t0) * 1e3
                info = plan.info()
                hot = plan.sddmm(K, dA, dB, dP, iterations=100)
                wms, dms, rms = [], [], []
                for _ in range(10):
                    flush.fill_(1)
                    w, a, b = plan.sddmm_profile3(K, dA, dB, dP)
                    wms.append(w)
                    dms.append(a)
                    rms.append(b)
                hot_nowide = plan.sddmm(K, dA, dB, dP, iterations=100, flags=pkg.SDDMM_NO_WIDE)
                tag = "d%.1f" % delta
                out[tag] = {"col_ms": info["col_reordering_ms"], "fmt_ms": info["format_build_ms"], "col_wall_ms": col_wall,
                            "dense_nnz": int(info["num_dense_values"]), "res_nnz": int(info["num_sparse_values"]),
                            "tiles": info["num_dense_tiles"], "hot_ms": hot, "hot_gflops": 2.0 * nnz * K / hot / 1e6,
                            "hot_nowide_ms": hot_nowide,
                            "wide": [info["num_wide_groups"], info["num_row_groups"], info["num_wide_tiles"], int(info["num_wide_values"]),
                                     int(info["num_block_values"]), int(info["num_residual_values"]), info["wide_format_ms"]],
                            "cold_wide_ms": float(np.median(wms)),
                            "cold_dense_ms": float(np.median(dms)), "cold_res_ms": float(np.median(rms)),
                            "cold_gflops": 2.0 * nnz * K / (np.median(wms) + np.median(dms) + np.median(rms)) / 1e6}
            hot = plan.sddmm(K, dA, dB, dP, iterations=100, flags=pkg.SDDMM_NO_REORDER)
            out["csr_order"] = {"hot_ms": hot, "hot_gflops": 2.0 * nnz * K / hot / 1e6}
            print(json.dumps(out), flush=True)
            del dA, dB, dP
        plan.close()


if __name__ == "__main__":
    main()
