"""B200 box (not a pytest test): row-reorder (BSA clustering) time against the row count on R-MAT graphs.

    python tests/cluster_scale_probe.py <scale> [edges_per_row=28.6] [alpha=0.3]

One JSON line: rows, nnz, block size, clusters, row-reorder ms (clustering kernel ms inside it).  Run one scale per
process under `timeout` -- the clustering pipeline's cost grows with clusters x rows.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    from graph8m_probe import rmat_device
    scale = int(sys.argv[1])
    per_row = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0e6 / (1 << 20)
    alpha = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    pkg = entry.load_package()
    edges = int(per_row * (1 << scale))
    n, ro, ci, rows = rmat_device(torch, scale, edges, seed=scale)
    del rows
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    plan = pkg.Plan(ctx, n, n, ro, ci, on_device=True)
    t0 = time.perf_counter()
    plan.row_reorder(alpha)
    wall = (time.perf_counter() - t0) * 1e3
    info = plan.info()
    print(json.dumps({"scale": scale, "rows": n, "nnz": edges, "alpha": alpha, "block_size": info["block_size"],
                      "clusters": info["num_clusters_true"], "row_reorder_ms": info["row_reordering_ms"],
                      "cluster_kernel_ms": info["cluster_kernel_ms"], "wall_ms": wall,
                      "nonempty_rows": int(len(plan.vector("reordered_rows")))}), flush=True)


if __name__ == "__main__":
    main()
