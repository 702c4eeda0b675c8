"""Diagnostic script for the B200 box (not a pytest test): exercises each stage of the hot path once
and prints what it finds, so that one `gpurun` call answers as many hardware questions as possible.

  python tests/gpu_probe.py [--skip-dense]
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle.bindings import Oracle, Ref, REF_SO  # noqa: E402


def section(title):
    print("\n==== " + title, flush=True)


def main():
    import torch
    pkg = entry.load_package()
    oracle = Oracle()
    print("device:", torch.cuda.get_device_name(0), "| lib:", pkg.lib().bsmr_version().decode())
    ctx = pkg.Context(0)
    M, N, ro, ci = pkg.synth.block_structured(1000, 2000, seed=11, groups=12, cols_per_group=64)
    nnz = len(ci)
    K = 128
    A, B = pkg.synth.make_ab(M, N, K)
    want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()

    section("CSR-order residual kernel")
    try:
        plan = pkg.Plan(ctx, M, N, ro, ci)
        dP = torch.zeros(nnz, device="cuda")
        ms = plan.sddmm(K, dA, dB, dP, iterations=5, flags=pkg.SDDMM_NO_REORDER)
        got = dP.cpu().numpy()
        print("ms/iter %.4f  mismatches %d  max rel %.3e" % (ms, oracle.check_data(want, got),
              np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3))))
    except Exception:
        traceback.print_exc()

    section("row reorder (GPU) vs oracle")
    try:
        t = time.time()
        plan = pkg.Plan(ctx, M, N, ro, ci)
        plan.row_reorder(0.3, block_size=16)
        got = plan.vector("reordered_rows")
        w, wc, wt = oracle.row_reordering(M, N, ro, ci, 0.3, 16)
        info = plan.info()
        print("equal:", np.array_equal(got, w), "clusters", info["num_clusters"], wc, info["num_clusters_true"], wt,
              "row_ms %.3f wall %.3f" % (info["row_reordering_ms"], time.time() - t))
        if not np.array_equal(got, w):
            d = np.nonzero(got[:min(len(got), len(w))] != w[:min(len(got), len(w))])[0]
            print("lens", len(got), len(w), "first diff at", d[:5], got[d[:5]], w[d[:5]])
    except Exception:
        traceback.print_exc()

    section("col reorder + format (GPU) vs oracle")
    try:
        rows = plan.vector("reordered_rows")
        for delta in (0.3, 0.0, 1.1):
            plan.col_reorder(delta)
            w = oracle.col_reordering(M, N, ro, ci, rows, delta, with_rphm=True)
            oks = {k: bool(np.array_equal(plan.vector(k), w[k])) for k in
                   ["dense_cols", "dense_col_offsets", "sparse_cols", "sparse_col_offsets", "sparse_value_offsets",
                    "block_values", "sparse_values", "sparse_relative_rows", "sparse_col_indices"]}
            info = plan.info()
            print("delta", delta, oks, "dense", info["num_dense_values"], "sparse", info["num_sparse_values"],
                  "tiles", info["num_dense_tiles"], "col_ms %.3f fmt_ms %.3f" % (info["col_reordering_ms"], info["format_build_ms"]))
    except Exception:
        traceback.print_exc()

    section("residual kernel in RPHM order (delta 1.1)")
    try:
        dP = torch.zeros(nnz, device="cuda")
        ms = plan.sddmm(K, dA, dB, dP, iterations=5)
        got = dP.cpu().numpy()
        print("ms/iter %.4f mismatches %d" % (ms, oracle.check_data(want, got)))
    except Exception:
        traceback.print_exc()

    if "--skip-dense" in sys.argv:
        return
    for mode in ("gather4",):
        section("dense tcgen05 kernel, TMA mode = " + mode)
        try:
            os.environ["BSMR_DENSE_TMA_MODE"] = mode
            plan.col_reorder(0.3)
            info = plan.info()
            dump = torch.zeros(16384 // 4 + 2048 // 4, dtype=torch.int32, device="cuda")
            pkg.lib().bsmr_debug_set_dense_smem_dump.argtypes = [__import__('ctypes').c_void_p]   # debug build only (make DEBUG=1, BSMR_B200_LIB)
            pkg.lib().bsmr_debug_set_dense_smem_dump(dump.data_ptr())
            dP = torch.full((nnz,), -7.0, device="cuda")
            ms = plan.sddmm(K, dA, dB, dP, iterations=1)
            torch.cuda.synchronize()
            got = dP.cpu().numpy()
            bad = oracle.check_data(want, got)
            rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
            print("ms %.4f mismatches %d / %d  unwritten %d  max rel %.3e mean rel %.3e" % (
                ms, bad, nnz, int(np.sum(got == -7.0)), rel.max(), rel.mean()))
            # split the error by path
            sv = plan.vector("sparse_values")
            is_sparse = np.zeros(nnz, dtype=bool)
            is_sparse[sv] = True
            if (~is_sparse).any():
                print("dense part: n %d max rel %.3e mean rel %.3e mean signed %.3e" % (
                    int((~is_sparse).sum()), rel[~is_sparse].max(), rel[~is_sparse].mean(),
                    float(np.mean((got[~is_sparse] - want[~is_sparse]) / want[~is_sparse]))))
            if is_sparse.any():
                print("residual part: n %d max rel %.3e" % (int(is_sparse.sum()), rel[is_sparse].max()))
            # smem image of the first tile's first stage
            img = dump.cpu().numpy().view(np.float32)
            bt = img[:4096].reshape(128, 32)
            at = img[4096:].reshape(16, 32)
            dcols = plan.vector("dense_cols")
            rrows = plan.vector("reordered_rows")
            ncols0 = min(128, int(plan.vector("dense_col_offsets")[1]))
            exp_b = np.zeros((128, 32), np.float32)
            for r in range(ncols0):
                if dcols[r] < N:
                    exp_b[r] = B[dcols[r], :32]
            exp_a = np.zeros((16, 32), np.float32)
            for r in range(min(16, len(rrows))):
                exp_a[r] = A[rrows[r], :32]

            def rna(x):   # cvt.rna.tf32.f32: round to nearest (ties away) on the low 13 mantissa bits
                b = x.view(np.uint32).astype(np.uint64)
                return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)
            exp_a, exp_b = rna(exp_a), rna(exp_b)

            def swz(x):
                out = np.zeros_like(x)
                for r in range(x.shape[0]):
                    for j in range(8):
                        out[r, ((j ^ (r % 8)) * 4):((j ^ (r % 8)) * 4 + 4)] = x[r, j * 4:j * 4 + 4]
                return out
            print("first tile ncols", ncols0, "| B tile plain match", np.array_equal(bt[:ncols0], exp_b[:ncols0]),
                  "swizzled match", np.array_equal(bt[:ncols0], swz(exp_b)[:ncols0]),
                  "| A tile plain", np.array_equal(at, exp_a), "swizzled", np.array_equal(at, swz(exp_a)))
            if not np.array_equal(bt[:ncols0], swz(exp_b)[:ncols0]):
                print("B row0 got", bt[0, :8], "exp", exp_b[0, :8])
                print("B row1 got", bt[1, :8], "exp(swz)", swz(exp_b)[1, :8])
                nzrows = np.nonzero(np.abs(bt).sum(axis=1))[0]
                print("non-zero smem rows:", nzrows[:40], "count", len(nzrows))
        except Exception:
            traceback.print_exc()
            break

    if os.path.exists(REF_SO) and "--with-ref" in sys.argv:
        section("reference GPU pipeline on the same input (oracle/_ref)")
        try:
            ref = Ref()
            out = ref.bsmr_sddmm_gpu(M, N, K, ro, ci, A, B, 0.3, 0.3, 16, iters=10)
            print("ref clusters", out["num_clusters"], "row_ms %.2f col_ms %.2f sddmm_ms %.4f" % (out["row_ms"], out["col_ms"], out["sddmm_ms"]),
                  "ref-vs-cpu mismatches", oracle.check_data(want, out["P"]),
                  "max rel %.3e" % np.max(np.abs(out["P"] - want) / np.maximum(np.abs(want), 1e-3)))
            print("rows equal to ours:", np.array_equal(out["reordered_rows"], plan.vector("reordered_rows")))
            _, cms = ref.cusparse_sddmm(M, N, K, ro, ci, A, B, iters=10)
            print("cusparse ms/iter %.4f" % cms)
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()
