"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- work-balanced contiguous panel ranges,
disjoint cover of the CSR value array, assembly of P by an all-reduce -- with the oracle standing in for the kernels
(no GPU in this container).  The same range arithmetic is what bsmr_plan_set_shard implements on the device side
(tests/test_gpu_parity.py::test_shards_partition_the_nnz checks that one on a B200)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_bounds(prefix, world):
    """Boundaries of bsmr_plan_set_shard (csrc/capi.cu): first panel whose nnz prefix reaches rank * total / world."""
    panels = len(prefix) - 1
    total = int(prefix[-1])
    out = [0]
    for r in range(1, world):
        target = total // world * r + (total % world) * r // world
        out.append(int(np.searchsorted(prefix, target, side="left")))
    out.append(panels)
    return [min(max(b, 0), panels) for b in out]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    from oracle.bindings import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    oracle = Oracle()
    M, N, ro, ci = pkg.synth.block_structured(600, 900, seed=21, groups=10, cols_per_group=48)
    K = 32
    A, B = pkg.synth.make_ab(M, N, K)
    # B is produced on rank 0 and broadcast, like bench.py does over NCCL
    tB = torch.from_numpy(B if rank == 0 else np.zeros_like(B))
    dist.broadcast(tB, src=0)
    B = tB.numpy()
    rows, _, _ = oracle.row_reordering(M, N, ro, ci, 0.3, 16)
    cr = oracle.col_reordering(M, N, ro, ci, rows, 0.3, with_rphm=True)
    panels = cr["num_row_panels"]
    row_nnz = np.diff(ro.astype(np.int64))[rows]
    panel_nnz = np.add.reduceat(row_nnz, np.arange(0, len(rows), 16))
    prefix = np.concatenate([[0], np.cumsum(panel_nnz)])
    bounds = shard_bounds(prefix, world)
    p0, p1 = bounds[rank], bounds[rank + 1]
    # this rank's nnz: the rows of its panel range
    mine = np.zeros(len(ci), dtype=bool)
    for r in rows[p0 * 16:p1 * 16]:
        mine[ro[r]:ro[r + 1]] = True
    full = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
    P = np.where(mine, full, 0).astype(np.float32)
    t = torch.from_numpy(P.copy())
    dist.all_reduce(t)                                   # shards are disjoint index sets: the sum assembles P
    cover = torch.from_numpy(mine.astype(np.int32))
    dist.all_reduce(cover)
    ok = bool(np.array_equal(t.numpy(), full)) and bool((cover.numpy() == 1).all())
    balance = float(mine.sum()) / (len(ci) / world)
    q.put((rank, ok, balance, int(p1 - p0), panels))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_sddmm_assembles_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert sum(r[3] for r in res) == res[0][4]
    for r in res:
        assert 0.7 < r[2] < 1.3, "nnz balance off: %s" % (res,)


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    nnz = rng.integers(1, 500, size=1000)
    prefix = np.concatenate([[0], np.cumsum(nnz)])
    for world in (1, 2, 3, 4, 8):
        b = shard_bounds(prefix, world)
        assert b[0] == 0 and b[-1] == 1000 and all(x <= y for x, y in zip(b, b[1:]))
        shares = [prefix[b[i + 1]] - prefix[b[i]] for i in range(world)]
        assert max(shares) - min(shares) <= 2 * nnz.max()


def test_work_prefix_balances_tiles_not_only_nnz():
    """bsmr_plan_set_shard balances on nnz + 3000 per wide tile (csrc/capi.cu): with row groups that differ 5x in nnz but
    not in tiles, the nnz prefix gives some rank several times the tiles of another, the work prefix does not."""
    rng = np.random.default_rng(1)
    groups, ppg, tiles_per_group = 48, 16, 97
    nnz_g = np.where(rng.random(groups) < 0.25, 400_000, 80_000) + rng.integers(0, 5000, groups)
    panel_nnz = np.repeat(nnz_g // ppg, ppg)
    panel_work = panel_nnz + 3000 * tiles_per_group // ppg
    for world in (2, 4, 8):
        cover = {}
        for name, per_panel in (("nnz", panel_nnz), ("work", panel_work)):
            prefix = np.concatenate([[0], np.cumsum(per_panel)])
            b = [(x + ppg // 2) // ppg * ppg for x in shard_bounds(prefix, world)]      # snapped to row-group boundaries
            b[0], b[-1] = 0, groups * ppg
            assert all(b[i] <= b[i + 1] for i in range(world)), (name, b)
            cover[name] = [(b[i + 1] - b[i]) // ppg for i in range(world)]              # row groups (= tiles / 97) per rank
            assert sum(cover[name]) == groups
        assert max(cover["work"]) - min(cover["work"]) <= max(cover["nnz"]) - min(cover["nnz"])
        assert max(cover["work"]) <= 1.5 * groups / world + 1
