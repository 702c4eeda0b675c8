"""CPU: the oracle (oracle/bsmr_oracle.c) against the reference's own CPU code compiled unmodified
(oracle/_ref/libbsmr_ref.so, only where it was built) and against the committed golden fixtures
that were generated from that library (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from cases import named_case, small_cases

COL_VECS = ["dense_cols", "dense_col_offsets", "sparse_cols", "sparse_col_offsets", "sparse_value_offsets"]


def nonempty_rows(ro):
    return np.nonzero(np.diff(ro.astype(np.int64)))[0].astype(np.uint32)


def test_make_data_matches_reference(oracle, ref):
    assert np.array_equal(oracle.make_data(5000), ref.make_data(50, 100))
    # column-major B draws the same stream (a fresh default-seeded engine per call, src/Matrix.cpp:131)
    assert np.array_equal(oracle.make_data(640), ref.make_data(64, 10, col_major=True))


def test_sddmm_cpu_matches_reference(pkg, oracle, ref):
    for name, M, N, ro, ci in small_cases(pkg):
        for K in (32, 40, 128):
            A, B = pkg.synth.make_ab(M, N, K)
            assert np.array_equal(oracle.sddmm_cpu(M, N, K, A, B, ro, ci), ref.sddmm_cpu(M, N, K, A, B, ro, ci)), (name, K)


@pytest.mark.parametrize("delta", [0.0, 0.1, 0.3, 0.5, 0.9, 1.1])
def test_col_reordering_matches_reference(pkg, oracle, ref, delta):
    for name, M, N, ro, ci in small_cases(pkg):
        rows = nonempty_rows(ro)
        rows = rows[np.random.default_rng(5).permutation(len(rows))]
        a = oracle.col_reordering(M, N, ro, ci, rows, delta)
        b = ref.col_reordering_cpu(M, N, ro, ci, rows, delta)
        for k in COL_VECS:
            assert np.array_equal(a[k], b[k]), (name, delta, k)


def test_check_data_matches_reference(oracle, ref):
    rng = np.random.default_rng(0)
    a = (rng.random(20000, dtype=np.float32) * 100).astype(np.float32)
    b = a * (1 + rng.normal(0, 7e-4, size=a.shape)).astype(np.float32)
    b[:100] = a[:100] + 5e-6
    a[100:200] = 0
    b[100:200] = rng.normal(0, 1e-5, 100).astype(np.float32)
    assert oracle.check_data(a, b) == ref.check_data(a, b)
    assert 0 < oracle.check_data(a, b) < len(a)


def test_mtx_loader_matches_reference(pkg, oracle, ref, tmp_path):
    M, N, ro, ci = pkg.synth.random_uniform(50, 70, 600, seed=8)
    vals = np.random.default_rng(1).random(len(ci)).astype(np.float32)
    p = str(tmp_path / "m.mtx")
    pkg.synth.write_mtx(p, M, N, ro, ci, values=vals, shuffle_seed=3)   # shuffled lines: row-stable sort only
    a = oracle.load_mtx(p)
    b = ref.load_matrix_file(p)
    assert a is not None and b is not None
    assert a[0] == b[0] and a[1] == b[1]
    for x, y in zip(a[2:], b[2:]):
        assert np.array_equal(x, y)
    # in-row column order is file order, not sorted
    assert not all(np.all(np.diff(a[3][a[2][r]:a[2][r + 1]].astype(np.int64)) > 0) for r in range(M))


def test_mtx_loader_rejections(oracle, ref, tmp_path):
    cases = {
        "dup.mtx": "%%MatrixMarket\n3 3 3\n1 1 1\n2 2 1\n1 1 2\n",
        "range.mtx": "%%MatrixMarket\n3 3 2\n1 1 1\n4 2 1\n",
        "few.mtx": "%%MatrixMarket\n3 3 3\n1 1 1\n2 2 1\n",
        "many.mtx": "%%MatrixMarket\n3 3 2\n1 1 1\n2 2 1\n3 3 1\n",
        "one.mtx": "%%MatrixMarket\n3 3 1\n1 1 1\n",
        "ok_novalue.mtx": "%%MatrixMarket\n% comment\n3 4 3\n3 1\n1 4\n\n2 2\n",
    }
    for name, text in cases.items():
        p = str(tmp_path / name)
        open(p, "w").write(text)
        a, b = oracle.load_mtx(p), ref.load_matrix_file(p)
        assert (a is None) == (b is None), name
        if a is not None:
            for x, y in zip(a, b):
                assert np.array_equal(x, y), name


def test_golden_fixtures(pkg, oracle, golden_dir):
    """Fixtures produced by the reference library (tests/golden/make_golden.py); they travel to the GPU box."""
    files = sorted(f for f in os.listdir(golden_dir) if f.endswith(".npz") and not f.endswith(("_perm_ref_gpu.npz", "_perm_oracle.npz")))
    assert files, "no golden fixtures committed"
    for f in files:
        g = np.load(os.path.join(golden_dir, f))
        M, N, K = int(g["M"]), int(g["N"]), int(g["K"])
        ro, ci = g["row_offsets"], g["col_indices"]
        A, B = pkg.synth.make_ab(M, N, K)
        assert np.array_equal(oracle.sddmm_cpu(M, N, K, A, B, ro, ci), g["P_ref_cpu"]), f
        rows = g["rows"]
        for delta in (0.1, 0.3):
            got = oracle.col_reordering(M, N, ro, ci, rows, delta)
            for k in COL_VECS:
                assert np.array_equal(got[k], g["d%02d_%s" % (int(delta * 10), k)]), (f, delta, k)
        if "perm_ref_gpu" in g.files:   # row permutation produced by the reference's GPU code on a B200
            perm, compat, _ = oracle.row_reordering(M, N, ro, ci, float(g["alpha"]), int(g["block_size"]))
            assert np.array_equal(perm, g["perm_ref_gpu"]), f
            assert compat == int(g["clusters_ref_gpu"]), f


def test_similarity_lossy_reduction_drops_warps(oracle):
    """7 reference warps (nb = 777 like nips): blocks owned by warps 2, 5 and 6 are ignored."""
    nb = 777
    bd = oracle.clustering_blockdim(nb)
    assert bd == 224
    rep = np.zeros(nb, dtype=np.uint32)
    cmp_ = np.zeros(nb, dtype=np.uint32)
    rep[10] = 3
    cmp_[10] = 3
    cmp_[70] = 9          # thread 70 -> warp 2 -> dropped
    assert oracle.similarity(rep, cmp_, bd) == 1.0
    assert oracle.similarity(rep, cmp_, bd, exact=True) < 1.0
    cmp2 = np.zeros(nb, dtype=np.uint32)
    cmp2[70] = 5          # only dropped blocks: lossy norm is 0 -> similarity 0
    assert oracle.similarity(rep, cmp2, bd) == 0.0


def test_row_reordering_invariants(pkg, oracle):
    """check_rowReordering (src/BSMR.cpp:444-486): no duplicates, no empty rows, nothing missing."""
    for name, M, N, ro, ci in small_cases(pkg):
        for alpha in (0.1, 0.5, 0.9):
            perm, compat, true = oracle.row_reordering(M, N, ro, ci, alpha, 16)
            assert sorted(perm.tolist()) == nonempty_rows(ro).tolist(), (name, alpha)
            assert true >= 1


def test_block_size_and_blockdim_rules(oracle):
    assert oracle.calculate_block_size(1500, 12419, 170 << 30) == 16
    assert oracle.calculate_block_size(1 << 20, 1 << 20, 170 << 30) == 171
    assert oracle.clustering_blockdim(10) == 32
    assert oracle.clustering_blockdim(256) == 64
    assert oracle.clustering_blockdim(777) == 224
    assert oracle.clustering_blockdim(6133) == 1024


def test_randomised_shapes_against_reference(pkg, oracle, ref):
    """40 seeded random shapes (tiny to a few thousand nnz, ragged panels, N not a multiple of 16, K not a multiple of 4):
    column reorder vectors and sddmm_cpu of the oracle are bit-identical to the reference's CPU code."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        M = int(rng.integers(1, 400))
        N = int(rng.integers(2, 700))
        nnz = int(rng.integers(2, max(3, min(M * N, 6000))))
        _, _, ro, ci = pkg.synth.random_uniform(M, N, nnz, seed=1000 + case)
        rows = nonempty_rows(ro)
        rows = rows[rng.permutation(len(rows))]
        delta = float(rng.choice([0.0, 0.05, 0.3, 0.6, 1.0, 1.1]))
        a = oracle.col_reordering(M, N, ro, ci, rows, delta)
        b = ref.col_reordering_cpu(M, N, ro, ci, rows, delta)
        for k in COL_VECS:
            assert np.array_equal(a[k], b[k]), (case, M, N, nnz, delta, k)
        K = int(rng.choice([1, 3, 32, 50, 64]))
        A, B = pkg.synth.make_ab(M, N, K, seed=case)
        assert np.array_equal(oracle.sddmm_cpu(M, N, K, A, B, ro, ci), ref.sddmm_cpu(M, N, K, A, B, ro, ci)), (case, K)


def test_indexed_restatement_equals_dense_restatement(pkg, oracle):
    """oracle_row_reordering_indexed (sparse encodings, rows filed per column block: the checker at graph scale) gives the
    permutation of oracle_row_reordering (the dense restatement pinned to the reference) -- with and without the filter
    that leaves the most popular blocks out, lossy and exact reduction."""
    for name, M, N, ro, ci in small_cases(pkg):
        for alpha in (0.1, 0.3, 0.7):
            for bs in (16, 37):
                for exact in (False, True):
                    want = oracle.row_reordering(M, N, ro, ci, alpha, bs, exact)
                    for filt in (True, False):
                        got = oracle.row_reordering_indexed(M, N, ro, ci, alpha, bs, exact, filt)
                        assert np.array_equal(got[0], want[0]) and got[1:] == want[1:], (name, alpha, bs, exact, filt)


def test_nips_permutation_golden(pkg, oracle, golden_dir):
    """The row permutation the reference's own GPU code produced on a B200 for the nips-shaped matrix (nb = 777 column
    blocks -> 7 warps -> the lossy reduction of include/cudaUtil.cuh:27-45), against the sparse restatement of the oracle."""
    path = os.path.join(golden_dir, "nips_perm_ref_gpu.npz")
    if not os.path.exists(path):
        pytest.skip("nips_perm_ref_gpu.npz not generated yet (tests/golden/make_golden.py --gpu on the B200 box)")
    g = np.load(path)
    M, N, ro, ci = pkg.synth.nips_like()
    assert int(g["input_checksum"]) == int(ci.astype(np.uint64).sum()) * 1000003 + int(ro.astype(np.uint64).sum())
    perm, compat, _ = oracle.row_reordering_indexed(M, N, ro, ci, float(g["alpha"]), int(g["block_size"]))
    assert np.array_equal(perm, g["perm_ref_gpu"]) and compat == int(g["num_clusters"])


def _kept_warps(blockdim, exact):
    """Which reference warps survive the shared-memory tree (include/cudaUtil.cuh:27-45): lossy for odd warp counts."""
    nw = blockdim // 32
    first = nw // 2
    if exact:
        p2 = 1
        while p2 < nw:
            p2 <<= 1
        first = p2 // 2
    contrib = [{w} if w < nw else set() for w in range(64)]
    s = first
    while s >= 1:
        for w in range(s):
            contrib[w] |= contrib[w + s]
        s >>= 1
    return contrib[0]


def _sparse_similarity(rep, cmp_, blockdim, exact):
    """The identity the stage clustering kernel decides with: max(a, c) = a + c - min(a, c) per block, so
    sim = min-sum / (L1(rep) + L1(row) - min-sum) over the blocks the reference's reduction keeps; min-sum only has terms where
    both are non-zero."""
    nb = len(rep)
    kept = np.isin((np.arange(nb) % blockdim) // 32, sorted(_kept_warps(blockdim, exact)))
    r = np.where(kept, rep, 0).astype(np.float64)
    c = np.where(kept, cmp_, 0).astype(np.float64)
    if not r.any() or not c.any():
        return None
    a, b = r / np.sqrt((r * r).sum()), c / np.sqrt((c * c).sum())
    m = np.minimum(a, b).sum()
    return m / (a.sum() + b.sum() - m)


@pytest.mark.parametrize("exact", [False, True])
def test_sparse_similarity_identity_against_the_oracle(oracle, exact):
    """The stage kernel takes a (row, representative) pair as decided when the sparse similarity is outside alpha +- 1e-3 and
    evaluates in the reference's operation order only inside: the two must agree far better than that.  Random sparse encodings,
    power-of-two and odd (lossy: 7, 21 warps) reference CTA sizes."""
    rng = np.random.default_rng(11)
    worst = 0.0
    for nb in (40, 777, 2048, 2532, 6132):
        bd = oracle.clustering_blockdim(nb)
        for _ in range(60):
            n_r, n_c = int(rng.integers(1, 200)), int(rng.integers(1, 200))
            rep = np.zeros(nb, dtype=np.uint32)
            cmp_ = np.zeros(nb, dtype=np.uint32)
            hubs = rng.integers(0, max(1, nb // 16), 8)                         # shared hub blocks, like an R-MAT graph
            for enc, n in ((rep, n_r), (cmp_, n_c)):
                idx = np.concatenate([rng.integers(0, nb, n), hubs[: rng.integers(0, 9)]])
                np.add.at(enc, idx, rng.integers(1, 40, len(idx)).astype(np.uint32))
            want = _sparse_similarity(rep, cmp_, bd, exact)
            if want is None:
                continue
            got = oracle.similarity(rep, cmp_, bd, exact=exact)
            worst = max(worst, abs(got - want))
    assert worst < 1e-6, worst            # measured: 7.6e-8


def test_single_block_clusters_normalise_to_exactly_one():
    """Certain joins of the stage kernel: a representative that is one column block with count c normalises to c / sqrt(c*c) in
    fp32, which is exactly 1 for every count below 2^16 (above, the reference's uint32 c*c wraps) -- so a single-run row joining
    it changes nothing that any later decision reads."""
    c = np.arange(1, 65536, dtype=np.uint64)
    nr = np.sqrt((c * c).astype(np.float32))
    assert np.all(c.astype(np.float32) / nr == np.float32(1.0))


def test_graph_goldens_are_permutations_with_consistent_counts(pkg, golden_dir):
    """The oracle-made graph fixtures (tests/golden/make_graph_golden.py): each holds a permutation of the graph's non-empty rows
    and cluster counts that fit it; the GPU tests compare the clustering kernels with them."""
    files = sorted(f for f in os.listdir(golden_dir) if f.endswith("_perm_oracle.npz"))
    assert files
    for f in files:
        g = np.load(os.path.join(golden_dir, f))
        scale = int(f[len("graph"):f.index("_")])
        _, M, N, ro, ci = named_case(pkg, "graph%d" % scale)
        assert (M, N, len(ci)) == (int(g["M"]), int(g["N"]), int(g["nnz"])), f
        nonempty = np.nonzero(np.diff(ro.astype(np.int64)))[0]
        assert np.array_equal(np.sort(g["perm"]), nonempty.astype(np.uint32)), f
        assert 1 < int(g["num_clusters_true"]) <= len(nonempty), f
