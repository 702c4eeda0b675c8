"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the parity checkers.

* ``Oracle``  -> oracle/_build/libbsmr_oracle.so (our C restatement, oracle/bsmr_oracle.c)
* ``Ref``     -> oracle/_ref/libbsmr_ref.so      (the unmodified reference compiled by oracle/Makefile)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
arms may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libbsmr_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbsmr_ref.so")

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)


def _p(a, t):
    return a.ctypes.data_as(t)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def build_ref():
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class _CSR(C.Structure):
    _fields_ = [("rows", C.c_uint32), ("cols", C.c_uint32), ("nnz", C.c_uint32),
                ("row_offsets", u32p), ("col_indices", u32p), ("values", f32p)]


class _ColReorder(C.Structure):
    _fields_ = [("num_row_panels", C.c_uint32),
                ("dense_cols", u32p), ("n_dense_cols", C.c_size_t),
                ("dense_col_offsets", u32p),
                ("sparse_cols", u32p), ("n_sparse_cols", C.c_size_t),
                ("sparse_col_offsets", u32p),
                ("sparse_value_offsets", u32p)]


class _Rphm(C.Structure):
    _fields_ = [("num_row_panels", C.c_uint32),
                ("block_offsets", u32p),
                ("block_values", u32p), ("n_block_values", C.c_size_t),
                ("sparse_values", u32p), ("sparse_relative_rows", u32p), ("sparse_col_indices", u32p),
                ("n_sparse", C.c_size_t)]


def _take(ptr, n):
    if n == 0:
        return np.zeros(0, dtype=np.uint32)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


class Oracle:
    """CPU restatement of the reference algorithm (see oracle/bsmr_oracle.h)."""

    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build_oracle()
        self.lib = lib = C.CDLL(path)
        lib.oracle_load_mtx.restype = C.c_int
        lib.oracle_load_mtx.argtypes = [C.c_char_p, C.POINTER(_CSR)]
        lib.oracle_free_csr.argtypes = [C.POINTER(_CSR)]
        lib.oracle_make_data.argtypes = [C.c_size_t, f32p]
        lib.oracle_sddmm_cpu.argtypes = [C.c_uint32] * 3 + [f32p, f32p, u32p, u32p, C.c_int, f32p]
        lib.oracle_check_one.restype = C.c_int
        lib.oracle_check_one.argtypes = [C.c_float, C.c_float]
        lib.oracle_check_data.restype = C.c_uint64
        lib.oracle_check_data.argtypes = [C.c_uint64, f32p, f32p]
        lib.oracle_calculate_block_size.restype = C.c_uint32
        lib.oracle_calculate_block_size.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        lib.oracle_dispersion.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, C.c_uint32, u32p, u32p]
        lib.oracle_clustering_blockdim.restype = C.c_uint32
        lib.oracle_clustering_blockdim.argtypes = [C.c_uint32]
        lib.oracle_similarity.restype = C.c_float
        lib.oracle_similarity.argtypes = [u32p, u32p, C.c_uint32, C.c_uint32, C.c_int]
        lib.oracle_row_reordering.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, C.c_float, C.c_uint32, C.c_int,
                                              u32p, u32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.oracle_row_reordering_indexed.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, C.c_float, C.c_uint32, C.c_int, C.c_int,
                                                      u32p, u32p, C.POINTER(C.c_int), C.POINTER(C.c_int), u32p,
                                                      C.POINTER(C.c_uint64)]
        lib.oracle_col_reordering.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p, C.c_uint32, C.c_float,
                                              C.POINTER(_ColReorder)]
        lib.oracle_free_colreorder.argtypes = [C.POINTER(_ColReorder)]
        lib.oracle_build_rphm.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p, C.c_uint32,
                                          C.POINTER(_ColReorder), C.POINTER(_Rphm)]
        lib.oracle_free_rphm.argtypes = [C.POINTER(_Rphm)]

    def load_mtx(self, path):
        csr = _CSR()
        if not self.lib.oracle_load_mtx(path.encode(), C.byref(csr)):
            return None
        out = (csr.rows, csr.cols, _take(csr.row_offsets, csr.rows + 1), _take(csr.col_indices, csr.nnz),
               np.ctypeslib.as_array(csr.values, shape=(csr.nnz,)).copy())
        self.lib.oracle_free_csr(C.byref(csr))
        return out

    def make_data(self, n):
        out = np.empty(n, dtype=np.float32)
        self.lib.oracle_make_data(n, _p(out, f32p))
        return out

    def sddmm_cpu(self, M, N, K, A, B, row_offsets, col_indices, num_threads=0):
        A, B, ro, ci = _f32(A), _f32(B), _u32(row_offsets), _u32(col_indices)
        P = np.zeros(len(ci), dtype=np.float32)
        self.lib.oracle_sddmm_cpu(M, N, K, _p(A, f32p), _p(B, f32p), _p(ro, u32p), _p(ci, u32p), num_threads,
                                  _p(P, f32p))
        return P

    def check_data(self, a, b):
        a, b = _f32(a), _f32(b)
        return int(self.lib.oracle_check_data(len(a), _p(a, f32p), _p(b, f32p)))

    def calculate_block_size(self, M, N, free_mem):
        return int(self.lib.oracle_calculate_block_size(M, N, free_mem))

    def clustering_blockdim(self, nb):
        return int(self.lib.oracle_clustering_blockdim(nb))

    def dispersion(self, M, N, row_offsets, col_indices, block_size):
        ro, ci = _u32(row_offsets), _u32(col_indices)
        nb = int(np.ceil(np.float32(N) / np.float32(block_size)))
        enc = np.zeros((M, nb), dtype=np.uint32)
        disp = np.zeros(M, dtype=np.uint32)
        self.lib.oracle_dispersion(M, N, _p(ro, u32p), _p(ci, u32p), block_size, _p(enc, u32p), _p(disp, u32p))
        return enc, disp

    def similarity(self, rep, cmp_, blockdim, exact=False):
        rep, cmp_ = _u32(rep), _u32(cmp_)
        return float(self.lib.oracle_similarity(_p(rep, u32p), _p(cmp_, u32p), len(rep), blockdim, int(exact)))

    def row_reordering(self, M, N, row_offsets, col_indices, alpha, block_size, exact=False):
        ro, ci = _u32(row_offsets), _u32(col_indices)
        perm = np.zeros(max(M, 1), dtype=np.uint32)
        n = C.c_uint32(0)
        cc, ct = C.c_int(0), C.c_int(0)
        self.lib.oracle_row_reordering(M, N, _p(ro, u32p), _p(ci, u32p), alpha, block_size, int(exact),
                                       _p(perm, u32p), C.byref(n), C.byref(cc), C.byref(ct))
        return perm[:n.value].copy(), cc.value, ct.value

    def row_reordering_indexed(self, M, N, row_offsets, col_indices, alpha, block_size, exact=False, use_filter=True,
                               with_stats=False):
        """Same permutation as row_reordering, on sparse encodings with the rows filed per column block (graph scale)."""
        ro, ci = _u32(row_offsets), _u32(col_indices)
        perm = np.zeros(max(M, 1), dtype=np.uint32)
        n = C.c_uint32(0)
        cc, ct = C.c_int(0), C.c_int(0)
        stats = (C.c_uint64 * 4)()
        self.lib.oracle_row_reordering_indexed(M, N, _p(ro, u32p), _p(ci, u32p), alpha, block_size, int(exact),
                                               int(use_filter), _p(perm, u32p), C.byref(n), C.byref(cc), C.byref(ct),
                                               None, stats)
        out = (perm[:n.value].copy(), cc.value, ct.value)
        if with_stats:
            out += (dict(evaluations=int(stats[0]), joins=int(stats[1]), filed=int(stats[2]), runs=int(stats[3])),)
        return out

    def col_reordering(self, M, N, row_offsets, col_indices, rows, delta, with_rphm=False):
        ro, ci, rows = _u32(row_offsets), _u32(col_indices), _u32(rows)
        cr = _ColReorder()
        self.lib.oracle_col_reordering(M, N, _p(ro, u32p), _p(ci, u32p), _p(rows, u32p), len(rows), delta,
                                       C.byref(cr))
        P = cr.num_row_panels
        out = dict(num_row_panels=P,
                   dense_cols=_take(cr.dense_cols, cr.n_dense_cols),
                   dense_col_offsets=_take(cr.dense_col_offsets, P + 1),
                   sparse_cols=_take(cr.sparse_cols, cr.n_sparse_cols),
                   sparse_col_offsets=_take(cr.sparse_col_offsets, P + 1),
                   sparse_value_offsets=_take(cr.sparse_value_offsets, P + 1))
        if with_rphm:
            r = _Rphm()
            self.lib.oracle_build_rphm(M, N, _p(ro, u32p), _p(ci, u32p), _p(rows, u32p), len(rows), C.byref(cr),
                                       C.byref(r))
            out.update(block_offsets=_take(r.block_offsets, P + 1),
                       block_values=_take(r.block_values, r.n_block_values),
                       sparse_values=_take(r.sparse_values, r.n_sparse),
                       sparse_relative_rows=_take(r.sparse_relative_rows, r.n_sparse),
                       sparse_col_indices=_take(r.sparse_col_indices, r.n_sparse))
            self.lib.oracle_free_rphm(C.byref(r))
        self.lib.oracle_free_colreorder(C.byref(cr))
        return out


SLOTS = dict(row_offsets=0, col_indices=1, reordered_rows=2, dense_cols=3, dense_col_offsets=4,
             sparse_cols=5, sparse_col_offsets=6, sparse_value_offsets=7)


class Ref:
    """The reference's own code (compiled unmodified) behind oracle/ref_harness.cu."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = lib = C.CDLL(path)
        lib.ref_result_size.restype = C.c_uint64
        lib.ref_result_size.argtypes = [C.c_int]
        lib.ref_result_copy.argtypes = [C.c_int, u32p]
        lib.ref_values_size.restype = C.c_uint64
        lib.ref_values_copy.argtypes = [f32p]
        lib.ref_load_matrix_file.restype = C.c_int
        lib.ref_load_matrix_file.argtypes = [C.c_char_p, u32p, u32p, u32p]
        lib.ref_make_data.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, f32p]
        lib.ref_sddmm_cpu.argtypes = [C.c_uint32] * 4 + [f32p, f32p, u32p, u32p, C.c_int, f32p]
        lib.ref_omp_max_threads.restype = C.c_int
        lib.ref_check_data.restype = C.c_uint64
        lib.ref_check_data.argtypes = [C.c_uint64, f32p, f32p]
        lib.ref_col_reordering_cpu.argtypes = [C.c_uint32] * 3 + [u32p, u32p, u32p, C.c_uint32, C.c_float]
        lib.ref_calculate_block_size.restype = C.c_uint32
        lib.ref_calculate_block_size.argtypes = [C.c_uint32, C.c_uint32]
        lib.ref_row_reordering_gpu.restype = C.c_int
        lib.ref_row_reordering_gpu.argtypes = [C.c_uint32] * 3 + [u32p, u32p, C.c_float, C.c_uint32, f32p]
        lib.ref_bsmr_sddmm_gpu.restype = C.c_int
        lib.ref_bsmr_sddmm_gpu.argtypes = [C.c_uint32] * 4 + [u32p, u32p, f32p, f32p, C.c_float, C.c_float,
                                                              C.c_uint32, C.c_int, f32p, f32p]
        lib.ref_cusparse_sddmm.restype = C.c_float
        lib.ref_cusparse_sddmm.argtypes = [C.c_uint32] * 4 + [u32p, u32p, f32p, f32p, C.c_int, f32p]

    def slot(self, name):
        s = SLOTS[name]
        n = int(self.lib.ref_result_size(s))
        out = np.zeros(n, dtype=np.uint32)
        if n:
            self.lib.ref_result_copy(s, _p(out, u32p))
        return out

    def load_matrix_file(self, path):
        r, c, z = C.c_uint32(), C.c_uint32(), C.c_uint32()
        if not self.lib.ref_load_matrix_file(path.encode(), C.byref(r), C.byref(c), C.byref(z)):
            return None
        vals = np.zeros(int(self.lib.ref_values_size()), dtype=np.float32)
        if len(vals):
            self.lib.ref_values_copy(_p(vals, f32p))
        return r.value, c.value, self.slot("row_offsets"), self.slot("col_indices"), vals

    def make_data(self, rows, cols, col_major=False, num_threads=1):
        out = np.empty(rows * cols, dtype=np.float32)
        self.lib.ref_make_data(rows, cols, int(col_major), num_threads, _p(out, f32p))
        return out

    def omp_max_threads(self):
        return int(self.lib.ref_omp_max_threads())

    def sddmm_cpu(self, M, N, K, A, B, row_offsets, col_indices, num_threads=0):
        A, B, ro, ci = _f32(A), _f32(B), _u32(row_offsets), _u32(col_indices)
        P = np.zeros(len(ci), dtype=np.float32)
        self.lib.ref_sddmm_cpu(M, N, K, len(ci), _p(A, f32p), _p(B, f32p), _p(ro, u32p), _p(ci, u32p),
                               num_threads, _p(P, f32p))
        return P

    def check_data(self, a, b):
        a, b = _f32(a), _f32(b)
        return int(self.lib.ref_check_data(len(a), _p(a, f32p), _p(b, f32p)))

    def col_reordering_cpu(self, M, N, row_offsets, col_indices, rows, delta):
        ro, ci, rows = _u32(row_offsets), _u32(col_indices), _u32(rows)
        self.lib.ref_col_reordering_cpu(M, N, len(ci), _p(ro, u32p), _p(ci, u32p), _p(rows, u32p), len(rows), delta)
        names = ["dense_cols", "dense_col_offsets", "sparse_cols", "sparse_col_offsets", "sparse_value_offsets"]
        out = {k: self.slot(k) for k in names}
        out["num_row_panels"] = len(out["dense_col_offsets"]) - 1
        return out

    # ---- GPU-only -------------------------------------------------------------------
    def calculate_block_size(self, M, N):
        return int(self.lib.ref_calculate_block_size(M, N))

    def row_reordering_gpu(self, M, N, row_offsets, col_indices, alpha, block_size):
        ro, ci = _u32(row_offsets), _u32(col_indices)
        t = C.c_float(0)
        nc = self.lib.ref_row_reordering_gpu(M, N, len(ci), _p(ro, u32p), _p(ci, u32p), alpha, block_size,
                                             C.byref(t))
        return self.slot("reordered_rows"), int(nc), t.value

    def bsmr_sddmm_gpu(self, M, N, K, row_offsets, col_indices, A, B, alpha, delta, block_size, iters=10):
        A, B, ro, ci = _f32(A), _f32(B), _u32(row_offsets), _u32(col_indices)
        P = np.zeros(len(ci), dtype=np.float32)
        times = np.zeros(3, dtype=np.float32)
        nc = self.lib.ref_bsmr_sddmm_gpu(M, N, len(ci), K, _p(ro, u32p), _p(ci, u32p), _p(A, f32p), _p(B, f32p),
                                         alpha, delta, block_size, iters, _p(P, f32p), _p(times, f32p))
        names = ["reordered_rows", "dense_cols", "dense_col_offsets", "sparse_cols", "sparse_col_offsets",
                 "sparse_value_offsets"]
        out = {k: self.slot(k) for k in names}
        out.update(P=P, num_clusters=int(nc), row_ms=float(times[0]), col_ms=float(times[1]),
                   sddmm_ms=float(times[2]))
        return out

    def cusparse_sddmm(self, M, N, K, row_offsets, col_indices, A, B, iters=10):
        A, B, ro, ci = _f32(A), _f32(B), _u32(row_offsets), _u32(col_indices)
        P = np.zeros(len(ci), dtype=np.float32)
        ms = self.lib.ref_cusparse_sddmm(M, N, len(ci), K, _p(ro, u32p), _p(ci, u32p), _p(A, f32p), _p(B, f32p),
                                         iters, _p(P, f32p))
        return P, float(ms)
