"""Seeded synthetic inputs of the shapes BASELINE.json names.

numpy only; every generator returns a CSR pattern ``(M, N, row_offsets u32, col_indices u32)``
with the columns of a row in ascending order.  ``make_ab`` produces A (M x K row-major) and
B (K x N column-major, i.e. an [N, K] array) uniform in [0, 2) like the reference's
``Matrix::makeData`` (src/Matrix.cpp:117-138) but from a counter-based generator so that the
streams are reproducible and independent of thread count.
"""
import numpy as np

NIPS_SHAPE = (1500, 12419, 746316)   # UCI NIPS bag-of-words: dataset/nips.mtx is not in the reference tree


def csr_from_rows_cols(M, N, rows, cols):
    """COO (any order, unique) -> CSR with ascending columns."""
    key = rows.astype(np.int64) * N + cols.astype(np.int64)
    key = np.unique(key)
    r = (key // N).astype(np.int64)
    c = (key % N).astype(np.uint32)
    ro = np.zeros(M + 1, dtype=np.int64)
    ro[1:] = np.bincount(r, minlength=M)
    return M, N, np.cumsum(ro).astype(np.uint32), c


def random_uniform(M, N, nnz, seed):
    """nnz distinct positions chosen uniformly (small parity cases)."""
    rng = np.random.default_rng(seed)
    nnz = min(nnz, M * N)
    flat = rng.choice(M * N, size=nnz, replace=False)
    return csr_from_rows_cols(M, N, flat // N, flat % N)


def nips_like(seed=1500, M=NIPS_SHAPE[0], N=NIPS_SHAPE[1], nnz=NIPS_SHAPE[2]):
    """Stand-in for dataset/nips.mtx: Zipf(1.0) column popularity, lognormal row lengths, exact nnz."""
    rng = np.random.default_rng(seed)
    length = rng.lognormal(mean=6.1, sigma=0.5, size=M)
    length = np.maximum(1, np.minimum(N, np.floor(length * nnz / length.sum()))).astype(np.int64)
    # fix the rounding so that the total is exactly nnz
    diff = int(nnz - length.sum())
    order = rng.permutation(M)
    i = 0
    while diff != 0:
        r = order[i % M]
        if diff > 0 and length[r] < N:
            length[r] += 1
            diff -= 1
        elif diff < 0 and length[r] > 1:
            length[r] -= 1
            diff += 1
        i += 1
    logw = -np.log(np.arange(1, N + 1, dtype=np.float64))       # Zipf s = 1
    colperm = rng.permutation(N)                                 # popular words are not the low ids
    ro = np.zeros(M + 1, dtype=np.int64)
    ro[1:] = np.cumsum(length)
    ci = np.empty(int(ro[-1]), dtype=np.uint32)
    for r in range(M):
        # weighted sampling without replacement: Gumbel top-k
        g = logw + rng.gumbel(size=N)
        top = np.argpartition(-g, int(length[r]) - 1)[: int(length[r])]
        ci[ro[r]:ro[r + 1]] = np.sort(colperm[top]).astype(np.uint32)
    return M, N, ro.astype(np.uint32), ci


def dlmc_mask(sparsity, n=4096, seed=None):
    """DLMC-style unstructured transformer mask: every row has exactly round((1-s)*n) columns."""
    seed = int(round(sparsity * 100)) if seed is None else seed
    rng = np.random.default_rng(seed)
    per_row = int(round((1.0 - sparsity) * n))
    ci = np.empty((n, per_row), dtype=np.uint32)
    for r in range(n):
        ci[r] = np.sort(rng.choice(n, size=per_row, replace=False)).astype(np.uint32)
    ro = (np.arange(n + 1, dtype=np.int64) * per_row).astype(np.uint32)
    return n, n, ro, ci.reshape(-1)


def rmat(scale, edges, seed, a=0.57, b=0.19, c=0.19):
    """R-MAT (a, b, c, d) directed graph, deduplicated, exactly `edges` edges, 2^scale vertices."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    keys = np.zeros(0, dtype=np.int64)
    want = edges
    while len(keys) < edges:
        m = int((want - len(keys)) * 1.25) + 1024
        src = np.zeros(m, dtype=np.int64)
        dst = np.zeros(m, dtype=np.int64)
        for _ in range(scale):
            u = rng.random(m)
            src = (src << 1) | (u >= a + b)
            dst = (dst << 1) | (((u >= a) & (u < a + b)) | (u >= a + b + c))
        keys = np.unique(np.concatenate([keys, src * n + dst]))
    if len(keys) > edges:
        keys = np.sort(rng.choice(keys, size=edges, replace=False))
    rows = keys // n
    ro = np.zeros(n + 1, dtype=np.int64)
    ro[1:] = np.bincount(rows, minlength=n)
    return n, n, np.cumsum(ro).astype(np.uint32), (keys % n).astype(np.uint32)


def block_structured(M, N, seed, groups=8, cols_per_group=48, fill=0.6, noise=0.01):
    """Rows drawn from a few column-support groups (shuffled): exercises clustering and dense blocks."""
    rng = np.random.default_rng(seed)
    supports = [rng.choice(N, size=min(cols_per_group, N), replace=False) for _ in range(groups)]
    rows, cols = [], []
    for r in range(M):
        g = rng.integers(groups)
        pick = supports[g][rng.random(len(supports[g])) < fill]
        extra = np.nonzero(rng.random(N) < noise)[0]
        cs = np.unique(np.concatenate([pick, extra]))
        if r % 17 == 5:          # some empty rows
            cs = cs[:0]
        rows.append(np.full(len(cs), r, dtype=np.int64))
        cols.append(cs.astype(np.int64))
    return csr_from_rows_cols(M, N, np.concatenate(rows), np.concatenate(cols))


def make_ab(M, N, K, seed=5489):
    """A [M, K] row-major and B stored column-major as an [N, K] array; fp32 uniform [0, 2)."""
    rng = np.random.Generator(np.random.Philox(seed))
    A = (rng.random((M, K), dtype=np.float32) * np.float32(2.0)).astype(np.float32)
    B = (rng.random((N, K), dtype=np.float32) * np.float32(2.0)).astype(np.float32)
    return A, B


def write_mtx(path, M, N, row_offsets, col_indices, values=None, shuffle_seed=None):
    """MatrixMarket coordinate file, 1-based; optional shuffled line order (the loader sorts by row only)."""
    rows = np.repeat(np.arange(M, dtype=np.int64), np.diff(row_offsets.astype(np.int64)))
    cols = col_indices.astype(np.int64)
    order = np.arange(len(cols))
    if shuffle_seed is not None:
        order = np.random.default_rng(shuffle_seed).permutation(len(cols))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% generated by bsmr-sddmm_b200/synth.py\n")
        f.write("%d %d %d\n" % (M, N, len(cols)))
        for i in order:
            if values is None:
                f.write("%d %d 1\n" % (rows[i] + 1, cols[i] + 1))
            else:
                f.write("%d %d %.9g\n" % (rows[i] + 1, cols[i] + 1, values[i]))


def read_mtx(path):
    """MatrixMarket coordinate file -> CSR pattern with the reference loader's ordering
    (1-based, '%' comment lines, stable sort by row only: src/Matrix.cpp:399-480)."""
    rows, cols = [], []
    header = None
    with open(path) as f:
        for line in f:
            if header is None:
                if line.startswith("%"):
                    continue
                header = [int(float(x)) for x in line.split()[:3]]
                continue
            w = line.split()
            if len(w) >= 2:
                rows.append(int(w[0]) - 1)
                cols.append(int(w[1]) - 1)
    M, N, nnz = header
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.uint32)
    if len(rows) != nnz:
        raise ValueError("%s: %d entries, header says %d" % (path, len(rows), nnz))
    order = np.argsort(rows, kind="stable")
    ro = np.zeros(M + 1, dtype=np.int64)
    np.add.at(ro, rows + 1, 1)
    return M, N, np.cumsum(ro).astype(np.uint32), cols[order]


def algorithmic_bytes(M, N, K, row_offsets, col_indices):
    """Compulsory traffic of one SDDMM: 4K(M_nz + N_nz) + 8 nnz + 4(M + 1)."""
    nnz = len(col_indices)
    m_nz = int(np.count_nonzero(np.diff(row_offsets.astype(np.int64))))
    n_nz = int(len(np.unique(col_indices)))
    return 4 * K * (m_nz + n_nz) + 8 * nnz + 4 * (M + 1)
