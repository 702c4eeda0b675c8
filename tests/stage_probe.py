"""B200 box (not a pytest test): the stage clustering kernel (32 clusters per CTA) against the oracle on the small cases,
against the reference's GPU permutation on nips, and against the cluster-per-CTA kernel on graphs, with timings.

    python tests/stage_probe.py small            every small case x alpha x block size x reduction mode against the oracle
    python tests/stage_probe.py nips             nips against the golden, mask98 against the cluster-per-CTA kernel
    python tests/stage_probe.py graph <scale> [both|one] [alpha] [edges] [out.npy]   R-MAT 2^scale rows: stage kernel (and, with `both`, the
                                                 cluster-per-CTA kernel + comparison); optionally saves the permutation

One JSON line per measurement.  Run under `timeout`.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from cases import named_case, small_cases  # noqa: E402


def timed_reorder(pkg, plan, alpha, block_size, flags):
    t0 = time.perf_counter()
    plan.row_reorder(alpha, block_size=block_size, flags=flags)
    wall = (time.perf_counter() - t0) * 1e3
    info = plan.info()
    return {"wall_ms": round(wall, 2), "cluster_kernel_ms": round(info["cluster_kernel_ms"], 2), "clusters": info["num_clusters_true"],
            "compat": info["num_clusters"], "block_size": info["block_size"]}


def main():
    import torch
    what = sys.argv[1]
    pkg = entry.load_package()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    if what == "small":
        from oracle.bindings import Oracle
        oracle = Oracle()
        bad = n = 0
        for alpha in (0.0, 0.1, 0.3, 0.5, 0.7, 0.9):
            for name, M, N, ro, ci in small_cases(pkg):
                for block_size in (16, 37):
                    for mode in (pkg.ROW_REFERENCE_COMPAT, pkg.ROW_EXACT_REDUCE):
                        want, want_compat, want_true = oracle.row_reordering(M, N, ro, ci, alpha, block_size, exact=(mode == pkg.ROW_EXACT_REDUCE))
                        plan = pkg.Plan(ctx, M, N, ro, ci)
                        plan.row_reorder(alpha, block_size=block_size, flags=mode | pkg.ROW_STAGE_ON)
                        info = plan.info()
                        ok = (np.array_equal(plan.vector("reordered_rows"), want) and info["num_clusters"] == want_compat
                              and info["num_clusters_true"] == want_true)
                        n += 1
                        if not ok:
                            bad += 1
                            print(json.dumps({"FAIL": name, "alpha": alpha, "block_size": block_size, "mode": mode,
                                              "clusters": [info["num_clusters"], info["num_clusters_true"]], "want": [want_compat, want_true]}), flush=True)
                        plan.close()
        print(json.dumps({"small_cases": n, "failed": bad}), flush=True)
    elif what == "nips":
        _, M, N, ro, ci = named_case(pkg, "nips")
        g = np.load(os.path.join(ROOT, "tests", "golden", "nips_perm_ref_gpu.npz"))
        plan = pkg.Plan(ctx, M, N, ro, ci)
        r = timed_reorder(pkg, plan, 0.3, 16, pkg.ROW_STAGE_ON)
        r.update(case="nips", equal_golden=bool(np.array_equal(plan.vector("reordered_rows"), g["perm_ref_gpu"]) and r["compat"] == int(g["num_clusters"])))
        print(json.dumps(r), flush=True)
        for case in ("mask98", "mask90"):
            _, M, N, ro, ci = named_case(pkg, case)
            plan = pkg.Plan(ctx, M, N, ro, ci)
            a = timed_reorder(pkg, plan, 0.3, 16, pkg.ROW_STAGE_OFF)
            pa = plan.vector("reordered_rows")
            b = timed_reorder(pkg, plan, 0.3, 16, pkg.ROW_STAGE_ON)
            pb = plan.vector("reordered_rows")
            print(json.dumps({"case": case, "cluster_per_cta": a, "stage": b, "equal": bool(np.array_equal(pa, pb) and a["clusters"] == b["clusters"] and a["compat"] == b["compat"])}), flush=True)
    elif what == "graph":
        from graph8m_probe import rmat_device
        scale = int(sys.argv[2])
        both = len(sys.argv) > 3 and sys.argv[3] == "both"
        alpha = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
        edges = int(sys.argv[5]) if len(sys.argv) > 5 else int(30.0e6 / (1 << 20) * (1 << scale))
        n, ro, ci, rows = rmat_device(torch, scale, edges, seed=scale)
        del rows
        plan = pkg.Plan(ctx, n, n, ro, ci, on_device=True)
        b = timed_reorder(pkg, plan, alpha, 0, pkg.ROW_STAGE_ON)
        pb = plan.vector("reordered_rows")
        out = {"scale": scale, "alpha": alpha, "nnz": edges, "nonempty_rows": int(len(pb)), "stage": b,
               "perm_is_permutation_of_nonempty_rows": bool(len(np.unique(pb)) == len(pb))}
        if len(sys.argv) > 6:                       # keep the order for later runs (bench.py loads it through bsmr_plan_set_row_order)
            np.save(sys.argv[6], pb)
        if both:
            a = timed_reorder(pkg, plan, alpha, 0, pkg.ROW_STAGE_OFF)
            pa = plan.vector("reordered_rows")
            out["cluster_per_cta"] = a
            out["equal"] = bool(np.array_equal(pa, pb) and a["clusters"] == b["clusters"] and a["compat"] == b["compat"])
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
