"""torchrun target (N >= 2 GPUs, not a pytest file): the NCCL data plane of the sharded SDDMM end to end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 \
        tests/sharded_nccl_check.py

Rank 0 computes the row order (the clustering is global), bsmr_plan_bcast_row_order installs it everywhere, every rank
runs the column reorder itself, takes its shard, and bsmr_sddmm_sharded_host / bsmr_sddmm_sharded assemble P on the
root, which checks it against the oracle.  torch.distributed (gloo) only carries the 128-byte NCCL id and the barriers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    pkg = entry.load_package()
    ident = [pkg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    ctx = pkg.Context(local)
    ctx.comm_init(ident[0], rank, world)
    from oracle.bindings import Oracle
    oracle = Oracle()
    ok = True
    for case, K, ratio in (("blocks", 64, 1.0), ("blocks", 256, 1.0), ("graph14", 128, None), ("graph14", 33, None)):
        if case == "blocks":
            M, N, ro, ci = pkg.synth.block_structured(1000, 2000, seed=11, groups=12, cols_per_group=64)
        else:
            M, N, ro, ci = pkg.synth.rmat(14, 400_000, 14)
        A, B = pkg.synth.make_ab(M, N, K)
        plan = pkg.Plan(ctx, M, N, ro, ci)
        if ratio is not None:
            plan.set_wide_ratio(ratio)
        if rank == 0:
            plan.row_reorder(0.3, block_size=16)
        plan.bcast_row_order(0)
        plan.col_reorder(0.3)
        rows = plan.vector("reordered_rows")
        every = [None] * world
        dist.all_gather_object(every, rows.tobytes())
        assert all(e == every[0] for e in every), "row order differs between ranks"
        p0, p1, shard_nnz = plan.set_shard(rank, world)
        sizes = [None] * world
        dist.all_gather_object(sizes, int(shard_nnz))
        assert sum(sizes) == len(ci), (sizes, len(ci))
        hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
        hP = torch.full((len(ci),), float("nan")).pin_memory() if rank == 0 else None
        t = plan.sddmm_sharded_host(K, hA, hB, hP)
        if rank == 0:
            want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
            bad = oracle.check_data(want, hP.numpy())
            print("host path %s K=%d: mismatches %d, shard nnz %s, times %s" % (case, K, bad, sizes, {k: round(v, 3) for k, v in t.items() if k.endswith("_ms")}), flush=True)
            ok = ok and bad == 0
        # device-resident form: A everywhere, B produced on rank 0 and replicated by the library's own broadcast
        dA = torch.from_numpy(A).cuda()
        dB = torch.from_numpy(B).cuda() if rank == 0 else torch.zeros((N, K), device="cuda")
        ctx.comm_bcast(dB, N * K * 4, 0)
        dP = torch.full((len(ci),), float("nan"), device="cuda") if rank == 0 else None
        plan.sddmm_sharded(K, dA, dB, dP)
        torch.cuda.synchronize()
        if rank == 0:
            bad = oracle.check_data(want, dP.cpu().numpy())
            print("device path %s K=%d: mismatches %d" % (case, K, bad), flush=True)
            ok = ok and bad == 0
        dist.barrier()
        plan.close()
    ctx.comm_destroy()
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    if rank == 0 and flag[0]:
        print("SHARDED_OK", flush=True)
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
