"""GPU tests added in round 2 (run on a B200 with `-m gpu`), all through the C ABI:

* the pattern checks of bsmr_plan_create / bsmr_plan_set_row_order (the reference's loaders reject the same inputs,
  src/Matrix.cpp:442-465)
* the explicit execution-plan measurement (bsmr_plan_autotune) -- a default call measures nothing
* operands without a tensor-core path on a plan WITH dense tiles (K = 33, 5, 7: the flat-list route)
* the batched entry point: one launch per kernel for the whole batch, bit-identical to single calls
* batchedMatrixTranspose, fp16 storage of B, the sharded data plane at world = 1 (pack / gather-v / un-permute)
* BASELINE.json's configurations at full size: nips (row permutation vs the reference's GPU code, the 7-warp lossy
  reduction), the DLMC masks, R-MAT graphs (size-independent checks)
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import named_case, small_cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torch_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def blocks_case(pkg):
    return pkg.synth.block_structured(1000, 2000, seed=11, groups=12, cols_per_group=64)


def three_kernel_case(pkg):
    """300 rows at 12 % density (a wide row group) on top of 1000 block-structured rows (dense 16 x 16 blocks + residual
    after clustering): the default plan runs all three kernels on it."""
    rng = np.random.default_rng(21)
    N = 3000
    top = []
    for r in range(300):
        n = max(1, int(rng.binomial(N, 0.12)))
        top.append(np.sort(rng.choice(N, size=n, replace=False)).astype(np.uint32))
    M2, _, ro2, ci2 = pkg.synth.block_structured(1000, N, seed=11, groups=12, cols_per_group=64, noise=0.002)
    ro = np.zeros(300 + M2 + 1, dtype=np.int64)
    ro[1:301] = np.cumsum([len(r) for r in top])
    ro[301:] = ro[300] + ro2[1:].astype(np.int64)
    return 300 + M2, N, ro.astype(np.uint32), np.concatenate(top + [ci2])


def rel_err(got, want):
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)))


# ------------------------------------------------------------------------------------ validation
def test_plan_create_rejects_what_the_loaders_reject(pkg, ctx):
    M, N, ro, ci = pkg.synth.random_uniform(64, 96, 900, seed=1)
    bad = ro.copy()
    bad[10], bad[11] = ro[11], ro[10]                        # offsets not monotone
    if bad[10] != bad[11]:
        with pytest.raises(pkg.BsmrError) as e:
            pkg.Plan(ctx, M, N, bad, ci)
        assert e.value.status == 1
    c2 = ci.copy()
    c2[17] = N                                               # column out of range
    with pytest.raises(pkg.BsmrError) as e:
        pkg.Plan(ctx, M, N, ro, c2)
    assert e.value.status == 1 and "column" in str(e.value)
    # the same coordinate twice, adjacent in a sorted row ...
    r = int(np.argmax(np.diff(ro.astype(np.int64)) >= 3))
    c3 = ci.copy()
    c3[ro[r] + 1] = c3[ro[r]]
    with pytest.raises(pkg.BsmrError) as e:
        pkg.Plan(ctx, M, N, ro, c3)
    assert e.value.status == 1 and "twice" in str(e.value)
    # ... and far apart in an unsorted row (file order is allowed, duplicates are not)
    c4 = ci.copy()
    seg = c4[ro[r]:ro[r + 1]].copy()
    seg = seg[::-1].copy()
    seg[-1] = seg[0]
    c4[ro[r]:ro[r + 1]] = seg
    with pytest.raises(pkg.BsmrError) as e:
        pkg.Plan(ctx, M, N, ro, c4)
    assert e.value.status == 1 and "twice" in str(e.value)
    # an unsorted row without duplicates is fine
    c5 = ci.copy()
    c5[ro[r]:ro[r + 1]] = ci[ro[r]:ro[r + 1]][::-1]
    pkg.Plan(ctx, M, N, ro, c5).close()
    plan = pkg.Plan(ctx, M, N, ro, ci)
    rows = np.nonzero(np.diff(ro.astype(np.int64)))[0].astype(np.uint32)
    rows[3] = rows[5]
    with pytest.raises(pkg.BsmrError) as e:
        plan.set_row_order(rows)
    assert e.value.status == 1 and "twice" in str(e.value)


# ------------------------------------------------------------------------------------ execution plans
def test_default_call_measures_nothing_and_autotune_is_explicit(pkg, ctx, oracle):
    import torch
    M, N, ro, ci = pkg.synth.nips_like()
    K = 128
    A, B = pkg.synth.make_ab(M, N, K)
    want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.reorder(0.3, 0.3, flags=pkg.ROW_IDENTITY)
    dA, dB = torch_dev(A), torch_dev(B)
    p_def = torch.full((len(ci),), -5.0, device="cuda")
    p_three = torch.full((len(ci),), -6.0, device="cuda")
    launches = ctx.launch_count()
    plan.sddmm(K, dA, dB, p_def, timed=False)
    per_call = ctx.launch_count() - launches
    assert 1 <= per_call <= 3                                  # one pass of the plan's kernels, no measurement passes
    assert plan.execution_choice(K) == pkg.SDDMM_DEFAULT
    plan.sddmm(K, dA, dB, p_three, flags=pkg.SDDMM_THREE_KERNEL, timed=False)
    torch.cuda.synchronize()
    assert torch.equal(p_def, p_three)                         # the default IS the three-kernel plan, bit for bit
    chosen = plan.autotune(K, dA, dB, p_def)
    assert chosen in (pkg.SDDMM_DEFAULT, pkg.SDDMM_NO_WIDE, pkg.SDDMM_NO_REORDER)
    assert plan.execution_choice(K) == chosen
    p_def.fill_(-5.0)
    plan.sddmm(K, dA, dB, p_def)
    torch.cuda.synchronize()
    assert oracle.check_data(want, p_def.cpu().numpy()) == 0
    plan.set_execution_choice(K, pkg.SDDMM_NO_WIDE)
    assert plan.execution_choice(K) == pkg.SDDMM_NO_WIDE
    plan.set_execution_choice(K, pkg.SDDMM_DEFAULT)
    assert plan.execution_choice(K) == pkg.SDDMM_DEFAULT


@pytest.mark.parametrize("K", [33, 5, 7, 130])
def test_k_without_tensor_core_path_on_a_plan_with_dense_tiles(pkg, ctx, oracle, K):
    """K % 4 != 0 has no TMA row stride: a plan WITH dense tiles and wide groups routes every nnz through the CUDA-core
    kernel in reordered-row order instead of failing."""
    import torch
    M, N, ro, ci = blocks_case(pkg)
    A, B = pkg.synth.make_ab(M, N, K)
    want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.set_wide_ratio(1.0)
    plan.reorder(0.3, 0.3, block_size=16)
    info = plan.info()
    assert info["num_dense_tiles"] > 0
    dA, dB = torch_dev(A), torch_dev(B)
    for flags in (pkg.SDDMM_DEFAULT, pkg.SDDMM_NO_WIDE):
        dP = torch.full((len(ci),), float("nan"), device="cuda")
        plan.sddmm(K, dA, dB, dP, flags=flags)
        torch.cuda.synchronize()
        got = dP.cpu().numpy()
        assert oracle.check_data(want, got) == 0 and rel_err(got, want) < 2e-5, (K, flags)
    # sharded: the flat list is cut at panel boundaries
    acc = torch.zeros(len(ci), device="cuda")
    for rank in range(3):
        plan.set_shard(rank, 3)
        p = torch.full((len(ci),), float("nan"), device="cuda")
        plan.sddmm(K, dA, dB, p)
        torch.cuda.synchronize()
        acc += torch.nan_to_num(p)
    assert oracle.check_data(want, acc.cpu().numpy()) == 0


# ------------------------------------------------------------------------------------ batch (f1)
@pytest.mark.parametrize("K,flags_name", [(64, "SDDMM_THREE_KERNEL"), (128, "SDDMM_THREE_KERNEL"), (256, "SDDMM_THREE_KERNEL"),
                                          (128, "SDDMM_NO_WIDE"), (32, "SDDMM_NO_REORDER"), (40, "SDDMM_DEFAULT"), (33, "SDDMM_DEFAULT")])
def test_batched_launch_is_bit_identical_to_single_calls(pkg, ctx, oracle, K, flags_name):
    """sddmm_gpu_batch: every kernel of the plan runs ONCE for the whole batch (wide kernel: the tile range walked per
    element; dense kernel: (element, tile) work items; residual kernel: gridDim.y) and produces the bits of the
    single calls; element 3 is also checked against the oracle."""
    import torch
    flags = getattr(pkg, flags_name)
    M, N, ro, ci = three_kernel_case(pkg)
    nb = 9
    rng = np.random.Generator(np.random.Philox(77))
    A = (rng.random((nb, M, K), dtype=np.float32) * 2).astype(np.float32)
    B = (rng.random((nb, N, K), dtype=np.float32) * 2).astype(np.float32)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.set_wide_ratio(8.0)                   # the 12 % rows qualify for the wide kernel, the block-structured rows do not
    plan.reorder(0.3, 0.3, block_size=16)
    info = plan.info()
    if flags == pkg.SDDMM_THREE_KERNEL and K % 32 == 0:
        assert info["num_wide_tiles"] > 0 and info["num_block_tiles"] > 0 and info["num_residual_values"] > 0, info
    dA, dB = torch_dev(A), torch_dev(B)
    nnz = len(ci)
    single = torch.full((nb, nnz), float("nan"), device="cuda")
    launches = ctx.launch_count()
    for b in range(nb):
        plan.sddmm(K, dA[b], dB[b], single[b], flags=flags, timed=False)
    per_single = (ctx.launch_count() - launches) // nb
    batched = torch.full((nb, nnz), float("nan"), device="cuda")
    launches = ctx.launch_count()
    plan.sddmm_batch(nb, K, dA, dB, batched, flags=flags)
    assert ctx.launch_count() - launches == per_single, "the batch must cost the launches of ONE pass"
    torch.cuda.synchronize()
    assert not bool(torch.isnan(batched).any())
    assert torch.equal(single, batched)
    want = oracle.sddmm_cpu(M, N, K, A[3], B[3], ro, ci)
    assert oracle.check_data(want, batched[3].cpu().numpy()) == 0


def test_batched_matrix_transpose(pkg, ctx):
    import torch
    for (w, h, nb) in ((64, 32, 3), (100, 37, 5), (1, 9, 2), (513, 130, 4)):
        x = torch.randn(nb, h, w, device="cuda")
        y = torch.empty(nb, w, h, device="cuda")
        ctx.batched_transpose(w, h, nb, x, y)
        torch.cuda.synchronize()
        assert torch.equal(y, x.transpose(1, 2).contiguous()), (w, h, nb)


# ------------------------------------------------------------------------------------ fp16 B (f4)
@pytest.mark.parametrize("K", [32, 64, 128, 256, 40, 20])
def test_sddmm_f16b_within_the_reference_tolerance(pkg, ctx, oracle, K):
    """B stored as fp16 (11 significant bits, like the TF32 tensor-core operands), A fp32, fp32 accumulation.
    Tolerance = the reference's checkData (include/checkData.hpp:21-30: |d| < 1e-5 or relative < 1e-3)."""
    import torch
    for name, M, N, ro, ci in small_cases(pkg):
        if name not in ("blocks_1000x2000", "uniform_200x333_ragged", "wide_33x9000"):
            continue
        A, B = pkg.synth.make_ab(M, N, K)
        want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
        plan = pkg.Plan(ctx, M, N, ro, ci)
        plan.reorder(0.3, 0.3, block_size=16)
        dA, dB = torch_dev(A), torch_dev(B)
        dBh = torch.empty((N, K), dtype=torch.float16, device="cuda")
        ctx.convert_f32_to_f16(dB, dBh, N * K)
        torch.cuda.synchronize()
        assert torch.equal(dBh, dB.half())                     # round to nearest even, like torch
        for flags in (pkg.SDDMM_DEFAULT, pkg.SDDMM_NO_REORDER):
            dP = torch.full((len(ci),), float("nan"), device="cuda")
            plan.sddmm_f16b(K, dA, dBh, dP, flags=flags)
            torch.cuda.synchronize()
            got = dP.cpu().numpy()
            assert oracle.check_data(want, got) == 0, (name, K, flags)
            assert rel_err(got, want) < 5e-4, (name, K, flags, rel_err(got, want))
        # values fp16 can hold exactly: the fp16 path then computes the bits of the fp32 CUDA-core kernel
        Bq = B.astype(np.float16).astype(np.float32)
        dBq = torch_dev(Bq)
        p32 = torch.zeros(len(ci), device="cuda")
        p16 = torch.zeros(len(ci), device="cuda")
        plan.sddmm(K, dA, dBq, p32, flags=pkg.SDDMM_NO_REORDER)
        plan.sddmm_f16b(K, dA, dBq.half(), p16, flags=pkg.SDDMM_NO_REORDER)
        torch.cuda.synchronize()
        assert float((p32 - p16).abs().max() / p32.abs().max()) < 2e-6, (name, K)    # same products, other summation tree (8 instead of 4 k per lane)
        plan.close()


# ------------------------------------------------------------------------------------ sharded data plane, world = 1
def test_sharded_data_plane_single_rank(pkg, ctx, oracle):
    """The NCCL data plane with a communicator of one rank: upload of the shard's A rows out of pinned host memory,
    B upload + all-gather, kernels, pack, gather-v (nothing to receive), un-permute, D2H -- against the oracle.
    The two-rank exchange runs in tests/sharded_nccl_check.py (torchrun, 2 GPUs)."""
    import torch
    M, N, ro, ci = blocks_case(pkg)
    K = 64
    A, B = pkg.synth.make_ab(M, N, K)
    want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
    ctx1 = pkg.Context(0)
    ctx1.comm_init(pkg.comm_unique_id(), 0, 1)
    plan = pkg.Plan(ctx1, M, N, ro, ci)
    plan.set_wide_ratio(1.0)
    plan.reorder(0.3, 0.3, block_size=16)
    plan.set_shard(0, 1)
    hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
    hP = torch.zeros(len(ci)).pin_memory()
    t = plan.sddmm_sharded_host(K, hA, hB, hP)
    assert oracle.check_data(want, hP.numpy()) == 0
    assert t["shard_nnz"] == len(ci) and t["d2h_bytes"] == 4 * len(ci)
    assert t["h2d_bytes"] == 4 * K * (len(plan.vector("reordered_rows")) + N)      # only the non-empty rows of A travel
    # pageable host memory: the whole of A goes up, same result
    hP2 = np.zeros(len(ci), dtype=np.float32)
    t2 = plan.sddmm_sharded_host(K, A, B, hP2)
    assert oracle.check_data(want, hP2) == 0 and t2["h2d_bytes"] == 4 * K * (M + N)
    # device-resident form
    dA, dB = torch_dev(A), torch_dev(B)
    dP = torch.full((len(ci),), float("nan"), device="cuda")
    plan.sddmm_sharded(K, dA, dB, dP)
    torch.cuda.synchronize()
    assert oracle.check_data(want, dP.cpu().numpy()) == 0
    plan.close()
    ctx1.comm_destroy()


def test_sharded_data_plane_two_ranks(pkg):
    """torchrun with two ranks when the box has two GPUs (bsmr_ctx_comm_init, row-order broadcast, B all-gather,
    gather-v of P to the root); skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "sharded_nccl_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert "SHARDED_OK" in p.stdout


# ------------------------------------------------------------------------------------ BASELINE configs at full size (row g)
def run_ref_child(tmp_path, case, K, alpha, delta, block_size, rows_only=False):
    out = str(tmp_path / ("ref_%s_%d_%s.npz" % (case, K, "rows" if rows_only else "full")))
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_gpu_child.py"), case, str(K), str(alpha), str(delta), str(block_size), out]
    if rows_only:
        cmd.append("rows-only")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    if p.returncode != 0 or not os.path.exists(out):
        return None
    return dict(np.load(out))


def test_nips_row_permutation_vs_reference_gpu_and_golden(pkg, ctx, oracle, ref, tmp_path, golden_dir):
    """configs[0]/[1] at full size: nb = 777 column blocks -> the reference's clustering CTA has 7 warps and its block
    reduction drops three of them (include/cudaUtil.cuh:27-45).  Our permutation and numClusters must
    equal what the reference's own bsa_rowReordering_gpu produces on this B200, and the committed golden."""
    _, M, N, ro, ci = named_case(pkg, "nips")
    want = run_ref_child(tmp_path, "nips", 32, 0.3, 0.3, 16, rows_only=True)
    assert want is not None, "the reference's row reordering did not run"
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.row_reorder(0.3, block_size=16)
    got = plan.vector("reordered_rows")
    assert np.array_equal(got, want["reordered_rows"])
    assert plan.info()["num_clusters"] == int(want["num_clusters"])
    gold = os.path.join(golden_dir, "nips_perm_ref_gpu.npz")
    if os.path.exists(gold):
        g = np.load(gold)
        assert np.array_equal(got, g["perm_ref_gpu"]) and plan.info()["num_clusters"] == int(g["num_clusters"])
    # the sparse restatement of the oracle (the checker used at graph scale) agrees as well
    perm, compat, true = oracle.row_reordering_indexed(M, N, ro, ci, 0.3, 16)
    assert np.array_equal(got, perm) and compat == plan.info()["num_clusters"]


@pytest.mark.parametrize("K", [32, 128, 256])
def test_nips_full_pipeline(pkg, ctx, oracle, ref, tmp_path, K):
    """configs[0] (K = 32) and configs[1] (K = 128, 256): reorder vectors vs the reference's GPU pipeline, values vs
    sddmm_cpu for every execution plan; the reference's own sddmm_gpu values are compared where it produces any
    (its K > 32 kernels build a shuffle mask with `1 << tId`, tId <= 255, and write nothing on sm_100)."""
    import torch
    _, M, N, ro, ci = named_case(pkg, "nips")
    A, B = pkg.synth.make_ab(M, N, K)
    cpu = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.reorder(0.3, 0.3, block_size=16)
    want = run_ref_child(tmp_path, "nips", K, 0.3, 0.3, 16)
    if want is not None:
        for k in ["reordered_rows", "dense_cols", "dense_col_offsets", "sparse_cols", "sparse_col_offsets", "sparse_value_offsets"]:
            assert np.array_equal(plan.vector(k), want[k]), (K, k)
    else:
        assert K != 32, "the reference's K <= 32 path is expected to run on sm_100"
    dA, dB = torch_dev(A), torch_dev(B)
    for flags in (pkg.SDDMM_THREE_KERNEL, pkg.SDDMM_NO_WIDE, pkg.SDDMM_NO_REORDER):
        dP = torch.full((len(ci),), float("nan"), device="cuda")
        plan.sddmm(K, dA, dB, dP, flags=flags)
        torch.cuda.synchronize()
        got = dP.cpu().numpy()
        assert oracle.check_data(cpu, got) == 0, (K, flags)
        if want is not None and oracle.check_data(cpu, want["P"]) == 0:
            assert oracle.check_data(want["P"], got) == 0, (K, flags)


@pytest.mark.parametrize("sparsity", [70, 90, 98])
def test_dlmc_masks_full_size(pkg, ctx, oracle, sparsity):
    """configs[2]: 4096 x 4096 unstructured masks, K = 64: BSMR reorder (bit-exact row order against the sparse oracle,
    column vectors against the oracle) and every execution plan against sddmm_cpu."""
    import torch
    K = 64
    _, M, N, ro, ci = named_case(pkg, "mask%d" % sparsity)
    A, B = pkg.synth.make_ab(M, N, K)
    cpu = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.reorder(0.3, 0.3, block_size=16)
    rows = plan.vector("reordered_rows")
    perm, compat, _ = oracle.row_reordering_indexed(M, N, ro, ci, 0.3, 16)
    assert np.array_equal(rows, perm) and plan.info()["num_clusters"] == compat
    want = oracle.col_reordering(M, N, ro, ci, rows, 0.3)
    for k in ["dense_cols", "dense_col_offsets", "sparse_cols", "sparse_col_offsets", "sparse_value_offsets"]:
        assert np.array_equal(plan.vector(k), want[k]), k
    dA, dB = torch_dev(A), torch_dev(B)
    for flags in (pkg.SDDMM_THREE_KERNEL, pkg.SDDMM_NO_WIDE, pkg.SDDMM_NO_REORDER):
        dP = torch.full((len(ci),), float("nan"), device="cuda")
        plan.sddmm(K, dA, dB, dP, flags=flags)
        torch.cuda.synchronize()
        assert oracle.check_data(cpu, dP.cpu().numpy()) == 0, (sparsity, flags)


@pytest.mark.parametrize("alpha", [0.0, 0.1, 0.3, 0.5, 0.7, 0.9, 1.0])
def test_row_reorder_both_clustering_steps_vs_oracle(pkg, ctx, oracle, alpha):
    """Three forms of the clustering: the cluster-per-CTA kernel with a warp per candidate (per-warp scratch; the default where
    rows are long), the same kernel with a thread per candidate for the cheap rejections, and the stage kernel (32 consecutive
    clusters per CTA; the default on graph-shaped inputs).  All must give the oracle's permutation on every small case, lossy and
    exact reduction."""
    for name, M, N, ro, ci in small_cases(pkg):
        for block_size in (16, 37):
            for mode in (pkg.ROW_REFERENCE_COMPAT, pkg.ROW_EXACT_REDUCE):
                want, want_compat, want_true = oracle.row_reordering(M, N, ro, ci, alpha, block_size, exact=(mode == pkg.ROW_EXACT_REDUCE))
                for step in (pkg.ROW_THREAD_PRUNE_ON, pkg.ROW_THREAD_PRUNE_OFF, pkg.ROW_STAGE_ON):
                    plan = pkg.Plan(ctx, M, N, ro, ci)
                    plan.row_reorder(alpha, block_size=block_size, flags=mode | step)
                    assert np.array_equal(plan.vector("reordered_rows"), want), (name, alpha, block_size, mode, step)
                    info = plan.info()
                    assert info["num_clusters"] == want_compat and info["num_clusters_true"] == want_true, (name, alpha, block_size, mode, step)
                    plan.close()


def test_nips_and_mask_clustering_with_the_thread_prune_step(pkg, ctx, oracle, golden_dir):
    """Full-size inputs through the step form that is NOT their default: nips against the reference's GPU permutation
    (golden), the 98 % mask (4095 clusters) against the sparse oracle."""
    _, M, N, ro, ci = named_case(pkg, "nips")
    plan = pkg.Plan(ctx, M, N, ro, ci)
    g = np.load(os.path.join(golden_dir, "nips_perm_ref_gpu.npz"))
    for step in (pkg.ROW_THREAD_PRUNE_ON, pkg.ROW_STAGE_ON):
        plan.row_reorder(0.3, block_size=16, flags=step)
        assert np.array_equal(plan.vector("reordered_rows"), g["perm_ref_gpu"]) and plan.info()["num_clusters"] == int(g["num_clusters"]), step
    _, M, N, ro, ci = named_case(pkg, "mask98")
    perm, compat, _ = oracle.row_reordering_indexed(M, N, ro, ci, 0.3, 16)
    for step in (pkg.ROW_THREAD_PRUNE_ON, pkg.ROW_THREAD_PRUNE_OFF, pkg.ROW_STAGE_ON):
        plan = pkg.Plan(ctx, M, N, ro, ci)
        plan.row_reorder(0.3, block_size=16, flags=step)
        assert np.array_equal(plan.vector("reordered_rows"), perm) and plan.info()["num_clusters"] == compat, step


def test_graph_clustering_vs_sparse_oracle(pkg, ctx, oracle):
    """8 192-row R-MAT graph (234 k nnz): the clustering pipeline against the sparse restatement of the oracle (itself pinned to the
    dense restatement and the reference on every small case).  Larger graphs: tests/cluster_scale_probe.py (timing)."""
    _, M, N, ro, ci = named_case(pkg, "graph13")
    bs = ctx.calculate_block_size(M, N)
    perm, compat, true = oracle.row_reordering_indexed(M, N, ro, ci, 0.3, bs)
    for step in (pkg.ROW_THREAD_PRUNE_ON, pkg.ROW_THREAD_PRUNE_OFF, pkg.ROW_STAGE_ON):
        plan = pkg.Plan(ctx, M, N, ro, ci)
        plan.row_reorder(0.3, block_size=bs, flags=step)
        assert np.array_equal(plan.vector("reordered_rows"), perm), step
        assert plan.info()["num_clusters"] == compat and plan.info()["num_clusters_true"] == true, step
        plan.close()


@pytest.mark.parametrize("scale,alpha", [(15, 0.3), (16, 0.1), (16, 0.6), (17, 0.3)])
def test_stage_kernel_equals_cluster_per_cta_kernel_on_graphs(pkg, ctx, scale, alpha):
    """R-MAT graphs beyond what the CPU oracle finishes: the stage kernel (32 clusters per CTA, the default there) and the
    cluster-per-CTA kernel (pinned to the oracle and the reference on everything smaller) must produce the same permutation and
    cluster counts -- hub rows with thousands of column blocks, tens of thousands of clusters, joins in every phase."""
    _, M, N, ro, ci = named_case(pkg, "graph%d" % scale)
    out = []
    for step in (pkg.ROW_STAGE_ON, pkg.ROW_STAGE_OFF):
        plan = pkg.Plan(ctx, M, N, ro, ci)
        plan.row_reorder(alpha, flags=step)
        info = plan.info()
        out.append((plan.vector("reordered_rows"), info["num_clusters"], info["num_clusters_true"]))
        plan.close()
    assert out[0][1:] == out[1][1:] and out[0][2] > 1000
    assert np.array_equal(out[0][0], out[1][0])


@pytest.mark.parametrize("scale", [15, 16, 17])
def test_graph_clustering_vs_cpu_oracle_golden(pkg, ctx, golden_dir, scale):
    """R-MAT graphs of 2^15 / 2^16 (/ 2^17 where its fixture exists) rows (15 725 / 30 875 clusters) against permutations the CPU oracle computed once
    (`tests/golden/make_graph_golden.py`: minutes to half an hour on one core, hence a committed fixture): both clustering
    kernels, permutation and both cluster counts."""
    path = os.path.join(golden_dir, "graph%d_perm_oracle.npz" % scale)
    if not os.path.exists(path):
        pytest.skip("no golden for graph%d" % scale)
    g = np.load(path)
    _, M, N, ro, ci = named_case(pkg, "graph%d" % scale)
    assert (M, N, len(ci)) == (int(g["M"]), int(g["N"]), int(g["nnz"]))
    for step in (pkg.ROW_STAGE_ON, pkg.ROW_STAGE_OFF):
        plan = pkg.Plan(ctx, M, N, ro, ci)
        plan.row_reorder(float(g["alpha"]), block_size=int(g["block_size"]), flags=step)
        info = plan.info()
        assert np.array_equal(plan.vector("reordered_rows"), g["perm"]), step
        assert info["num_clusters"] == int(g["num_clusters"]) and info["num_clusters_true"] == int(g["num_clusters_true"]), step
        plan.close()


def test_stage_kernel_request_falls_back_when_it_does_not_fit(pkg, ctx):
    """18 750 column blocks do not fit the stage kernel's shared memory: BSMR_ROW_STAGE_ON must fall back to the cluster-per-CTA
    kernel (same permutation as BSMR_ROW_STAGE_OFF), not fail."""
    rng = np.random.default_rng(9)
    M, N = 1500, 300000
    keys = np.unique(rng.integers(0, M, 30000).astype(np.int64) * N + rng.integers(0, N, 30000))
    M, N, ro, ci = pkg.synth.csr_from_rows_cols(M, N, keys // N, keys % N)
    out = []
    for step in (pkg.ROW_STAGE_ON, pkg.ROW_STAGE_OFF):
        plan = pkg.Plan(ctx, M, N, ro, ci)
        plan.row_reorder(0.3, block_size=16, flags=step)
        out.append((plan.vector("reordered_rows"), plan.info()["num_clusters_true"]))
        plan.close()
    assert out[0][1] == out[1][1] and np.array_equal(out[0][0], out[1][0])


def sampled_check(torch, dA, dB, dP, ro_dev, ci_dev, nnz, n=1 << 17, seed=3):
    """size-independent parity: n sampled entries against fp64 dot products of the same rows"""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    idx = torch.randint(0, nnz, (n,), device="cuda", generator=g)
    rows = torch.searchsorted(ro_dev.long(), idx, right=True) - 1
    ref = (dA[rows].double() * dB[ci_dev[idx].long()].double()).sum(-1)
    return float(((dP[idx].double() - ref).abs() / ref.abs().clamp_min(1e-3)).max())


@pytest.mark.parametrize("scale,edges,K", [(20, 30_000_000, 128), (23, 250_000_000, 256)])
def test_graphs_full_size_properties(pkg, ctx, scale, edges, K):
    """configs[3] (2^20 rows, 3e7 nnz, K = 128) and configs[4] (2^23 rows, 2.5e8 nnz, K = 256) at full size, identity row
    order (the clustering at these sizes is timed by bench.py / tests/cluster_scale_probe.py, not here): the split covers
    every nnz, every entry is written, 2^17 sampled entries agree with fp64 dot products, fp32 and fp16-B paths."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from graph8m_probe import rmat_device
    free, _ = torch.cuda.mem_get_info()
    need = (2 * (1 << scale) * K * 4) * 1.6 + edges * 60
    if free < need:
        pytest.skip("needs %.0f GB of device memory" % (need / 1e9))
    n, ro, ci, rows = rmat_device(torch, scale, edges, seed=scale)
    del rows
    plan = pkg.Plan(ctx, n, n, ro, ci, on_device=True)
    plan.reorder(0.3, 0.3, flags=pkg.ROW_IDENTITY)
    info = plan.info()
    assert info["num_dense_values"] + info["num_sparse_values"] == edges
    assert info["num_wide_values"] + info["num_block_values"] + info["num_residual_values"] == edges
    g = torch.Generator(device="cuda")
    g.manual_seed(5489)
    dA = torch.rand((n, K), device="cuda", generator=g) * 2
    dB = torch.rand((n, K), device="cuda", generator=g) * 2
    dP = torch.full((edges,), float("nan"), device="cuda")
    plan.sddmm(K, dA, dB, dP)
    torch.cuda.synchronize()
    assert not bool(torch.isnan(dP).any()), "every entry must be written"
    assert sampled_check(torch, dA, dB, dP, ro, ci, edges) < 1e-3
    dBh = dB.half()
    dP.fill_(float("nan"))
    plan.sddmm_f16b(K, dA, dBh, dP)
    torch.cuda.synchronize()
    assert not bool(torch.isnan(dP).any())
    assert sampled_check(torch, dA, dB, dP, ro, ci, edges) < 1e-3
