"""B200 box: BASELINE.json configs[4] on one GPU (not a pytest test).

    python tests/graph8m_probe.py [scale=23] [edges=250000000] [K=256]

R-MAT (a, b, c, d = .57, .19, .19, .05) graph of 2^scale vertices, deduplicated, exactly `edges` edges, generated ON
THE GPU with torch (the numpy generator of `synth.rmat` needs minutes of host time at this size), handed to
`bsmr_plan_create` as device CSR; identity row order (million-row clustering is not built),
column reorder + format build at delta = 0.3, then `bsmr_sddmm` at K.  Prints one JSON line: reorder / format times, the
split, ms per SDDMM, GFLOPS, the fraction of the HBM roofline on SURVEY section 8(d)'s algorithmic bytes, and a sampled
check of the values against an fp64 dot product of the same rows (size-independent parity at the full size).

Under torchrun (`python -m torch.distributed.run --nproc-per-node N ... tests/graph8m_probe.py ...`) every rank builds
the same plan (same seed), takes the rank-th of N work-balanced shards of row panels (`bsmr_plan_set_shard`) and times
its own shard; rank 0 prints the line with the per-rank times and GFLOPS over the slowest rank (strong scaling: the
matrix is fixed).  A and B are replicated and there is no collective in the timed SDDMM; afterwards P is assembled on
every rank by one NCCL all-reduce of the disjoint shards, timed separately (`assemble_p_nccl_allreduce_ms`) and checked
on entries sampled from the whole matrix (the assembly step was added after the r01h runs of profiles/ and has not run
on a GPU box yet).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def rmat_device(torch, scale, edges, seed, a=0.57, b=0.19, c=0.19):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    keys = torch.zeros(0, dtype=torch.int64, device="cuda")
    while keys.numel() < edges:
        m = int((edges - keys.numel()) * 1.3) + 1024
        src = torch.zeros(m, dtype=torch.int64, device="cuda")
        dst = torch.zeros(m, dtype=torch.int64, device="cuda")
        for _ in range(scale):
            u = torch.rand(m, device="cuda", generator=g)
            src = (src << 1) | (u >= a + b)
            dst = (dst << 1) | (((u >= a) & (u < a + b)) | (u >= a + b + c))
            del u
        keys = torch.unique(torch.cat([keys, (src << scale) | dst]))
        del src, dst
    if keys.numel() > edges:
        keep = torch.randperm(keys.numel(), device="cuda", generator=g)[:edges]
        keys = keys[torch.sort(keep).values]
    n = 1 << scale
    rows = keys >> scale
    counts = torch.bincount(rows, minlength=n)
    ro = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    ro[1:] = torch.cumsum(counts, 0)
    ci = (keys & (n - 1)).to(torch.int32)
    return n, ro.to(torch.int32), ci, rows


def sharded(torch, pkg, plan, out, world, rank, K, dA, dB, dP, ro, ci, rows, g):
    import torch.distributed as dist
    nnz = ci.numel()
    first_panel, end_panel, shard_nnz = plan.set_shard(rank, world)
    plan.sddmm(K, dA, dB, dP, iterations=1)                    # execution-plan choice on the shard
    dP.zero_()
    dist.barrier()
    torch.cuda.synchronize()
    ms = min(plan.sddmm(K, dA, dB, dP, iterations=3) for _ in range(3))
    # the shard's entries: identity row order -> reordered row i is the i-th non-empty row, 16 rows per panel
    nz_rows = torch.nonzero(ro[1:] != ro[:-1]).flatten()
    r0 = int(nz_rows[first_panel * 16])
    r1 = int(nz_rows[min(end_panel * 16, nz_rows.numel()) - 1]) + 1
    e0, e1 = int(ro[r0]), int(ro[r1])
    idx = torch.randint(e0, e1, (1 << 17,), device="cuda", generator=g)
    ref = (dA[rows[idx]].double() * dB[ci[idx].long()].double()).sum(-1)
    rel = ((dP[idx].double() - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()
    written = int((dP != 0).sum())
    ok = rel <= 1e-3 and written == shard_nnz == e1 - e0
    # assembly of P over NCCL: the shards are disjoint index sets of the CSR value array and dP is zero elsewhere, so an
    # all-reduce(sum) leaves the complete result on every rank (timed on the device, outside the SDDMM time above)
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    dist.all_reduce(dP)
    ev1.record()
    torch.cuda.synchronize()
    assemble_ms = ev0.elapsed_time(ev1)
    idx_all = torch.randint(0, nnz, (1 << 17,), device="cuda", generator=g)
    ref_all = (dA[rows[idx_all]].double() * dB[ci[idx_all].long()].double()).sum(-1)
    rel_all = ((dP[idx_all].double() - ref_all).abs() / ref_all.abs().clamp_min(1e-30)).max().item()
    ok = ok and rel_all <= 1e-3 and int((dP == 0).sum()) == 0
    rel = max(rel, rel_all)
    mine = torch.tensor([ms, float(shard_nnz), float(end_panel - first_panel), rel, float(ok), assemble_ms], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    if rank == 0:
        t = max(float(a[0]) for a in allr)
        out.update({"n_gpus": world, "scaling": "strong", "ms_by_rank": [float(a[0]) for a in allr],
                    "shard_nnz": [int(a[1]) for a in allr], "shard_panels": [int(a[2]) for a in allr],
                    "sample_max_rel_err": max(float(a[3]) for a in allr), "parity_ok": all(float(a[4]) == 1.0 for a in allr),
                    "covered_nnz": sum(int(a[1]) for a in allr), "sddmm_ms": t, "gflops": 2.0 * nnz * K / t / 1e6,
                    "assemble_p_nccl_allreduce_ms": max(float(a[5]) for a in allr)})
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    import torch
    pkg = entry.load_package()
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    scale = int(args[0]) if len(args) > 0 else 23
    edges = int(args[1]) if len(args) > 1 else 250_000_000
    K = int(args[2]) if len(args) > 2 else 256
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    bw = float(peaks.get("hbm_gbs", 6552.3))

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(local, stream.cuda_stream)
    t0 = time.perf_counter()
    n, ro, ci, rows = rmat_device(torch, scale, edges, seed=scale)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    nnz = ci.numel()
    m_nz = int((ro[1:] != ro[:-1]).sum())
    n_nz = int(torch.unique(ci).numel())
    out = {"workload": "rmat%d" % scale, "M": n, "N": n, "nnz": nnz, "K": K, "M_nz": m_nz, "N_nz": n_nz, "generate_s": gen_s}

    t0 = time.perf_counter()
    plan = pkg.Plan(ctx, n, n, ro, ci, on_device=True)
    out["plan_create_ms"] = (time.perf_counter() - t0) * 1e3
    plan.row_reorder(0.3, flags=pkg.ROW_IDENTITY)
    for rep in range(2):   # second call: scratch arena warm
        t0 = time.perf_counter()
        plan.col_reorder(0.3)
        out["col_wall_ms" if rep else "col_wall_first_ms"] = (time.perf_counter() - t0) * 1e3
    info = plan.info()
    out.update({k: info[k] for k in ("num_row_panels", "col_reordering_ms", "format_build_ms", "num_dense_tiles",
                                     "num_row_groups", "num_wide_groups", "num_wide_tiles")})
    out.update({k: int(info[k]) for k in ("num_dense_values", "num_sparse_values", "num_wide_values", "num_block_values",
                                          "num_residual_values")})
    if rank == 0:
        print(json.dumps(out), file=sys.stderr, flush=True)

    g = torch.Generator(device="cuda")
    g.manual_seed(5489)
    dA = torch.rand((n, K), device="cuda", generator=g) * 2.0
    dB = torch.rand((n, K), device="cuda", generator=g) * 2.0
    dP = torch.zeros(nnz, device="cuda")
    torch.cuda.synchronize()
    out["hbm_used_gb"] = torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9

    if world > 1:
        sharded(torch, pkg, plan, out, world, rank, K, dA, dB, dP, ro, ci, rows, g)
        plan.close()
        return
    first = plan.sddmm(K, dA, dB, dP, iterations=1)           # includes the execution-plan choice
    ms = min(plan.sddmm(K, dA, dB, dP, iterations=3) for _ in range(3))
    out["first_call_ms"] = first
    out["execution_choice"] = plan.execution_choice(K)
    bytes_alg = 4.0 * K * (m_nz + n_nz) + 8.0 * nnz + 4.0 * (n + 1)
    out.update({"sddmm_ms": ms, "gflops": 2.0 * nnz * K / ms / 1e6, "bytes_alg_gb": bytes_alg / 1e9,
                "hbm_gbs_alg": bytes_alg / ms / 1e6, "hbm_frac": bytes_alg / ms / 1e6 / bw, "hbm_peak_gbs": bw,
                "gflops_with_reorder": 2.0 * nnz * K / (ms + info["col_reordering_ms"] + info["format_build_ms"]) / 1e6})

    # sampled parity at the full size: fp64 dot products of 2^18 random entries, plus "every entry was written"
    idx = torch.randint(0, nnz, (1 << 18,), device="cuda", generator=g)
    ref = (dA[rows[idx]].double() * dB[ci[idx].long()].double()).sum(-1)
    got = dP[idx].double()
    rel = ((got - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()
    out["sample_entries"] = int(idx.numel())
    out["sample_max_rel_err"] = rel
    out["entries_not_written"] = int((dP == 0).sum())
    out["parity_ok"] = bool(rel <= 1e-3 and out["entries_not_written"] == 0)

    ms_csr = min(plan.sddmm(K, dA, dB, dP, iterations=3, flags=pkg.SDDMM_NO_REORDER) for _ in range(2))
    got = dP[idx].double()
    out["csr_order_ms"] = ms_csr
    out["csr_order_max_rel_err"] = ((got - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()

    # L2 policy of the residual kernel (bsmr_plan_set_l2_policy): hub-column budget in MiB x priority of the other columns
    sweep = []
    keep = dP.clone()
    for budget, cold_first in ((0, 0), (32, 0), (64, 0), (96, 0), (32, 1), (64, 1), (96, 1)):
        plan.set_l2_policy(budget, 96, bool(cold_first))     # 96 MiB floor: the sweep also covers B = 537 MB
        dP.zero_()
        csr = min(plan.sddmm(K, dA, dB, dP, iterations=3, flags=pkg.SDDMM_NO_REORDER) for _ in range(2))
        same = bool(torch.equal(dP, keep))
        split = min(plan.sddmm(K, dA, dB, dP, iterations=3, flags=pkg.SDDMM_NO_WIDE) for _ in range(2))
        sweep.append({"hot_mb": budget, "cold_first": cold_first, "csr_order_ms": csr, "bsmr_split_ms": split, "bit_identical": same})
    out["l2_policy_sweep"] = sweep
    plan.set_l2_policy()
    print(json.dumps(out), flush=True)
    plan.close()


if __name__ == "__main__":
    main()
