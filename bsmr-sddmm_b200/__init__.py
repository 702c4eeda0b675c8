"""bsmr-sddmm_b200: B200 (sm_100a) implementation of the BSMR-SDDMM hot path.

The product is ``lib/libbsmr_b200.so`` (CUDA + a C ABI, see ``include/bsmr_b200.h``).  This
package is a thin ctypes binding over that ABI plus seeded synthetic inputs; it holds no
compute of its own and has no CPU fallback: if the shared library is missing, or no sm_100
device is present, every compute entry point raises.

The directory name carries a hyphen (it mirrors the reference's name), so it is imported as
``bsmr_sddmm_b200`` through ``importlib`` -- see ``load_package()`` in ``__graft_entry__.py``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "lib", "libbsmr_b200.so")
HEADER = os.path.join(ROOT, "include", "bsmr_b200.h")

ROW_PANEL_SIZE = 16
BLOCK_COL_SIZE = 16
NULL_VALUE = 0xFFFFFFFF

ROW_REFERENCE_COMPAT, ROW_EXACT_REDUCE, ROW_IDENTITY = 0, 1, 2
ROW_THREAD_PRUNE_ON, ROW_THREAD_PRUNE_OFF = 4, 8      # or-ed in: which step form the clustering kernel runs (default: by shape)
ROW_STAGE_ON, ROW_STAGE_OFF = 16, 32                   # or-ed in: the 32-clusters-per-CTA stage kernel or the cluster-per-CTA kernel (default: by shape)
SDDMM_DEFAULT, SDDMM_RESIDUAL_ONLY, SDDMM_NO_REORDER, SDDMM_NO_WIDE, SDDMM_THREE_KERNEL = 0, 1, 2, 4, 8
TICKET_ALL = 0xFFFFFFFFFFFFFFFF
CUDA_STREAM_LEGACY = 1          # cudaStreamLegacy: the legacy default stream as an explicit handle
WIDE_GROUP_ROWS, WIDE_TILE_COLS = 256, 128
WIDE_EPILOGUE_AUTO, WIDE_EPILOGUE_LIST, WIDE_EPILOGUE_MASK = 0, 1, 2

VEC = dict(reordered_rows=0, dense_cols=1, dense_col_offsets=2, sparse_cols=3, sparse_col_offsets=4,
           sparse_value_offsets=5, block_offsets=6, block_values=7, sparse_values=8,
           sparse_relative_rows=9, sparse_col_indices=10, dispersions=11, cluster_ids=12, group_wide=13)

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)


class BsmrError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("bsmr_b200 status %d: %s" % (status, message))
        self.status = status


class PlanInfo(C.Structure):
    _fields_ = [("M", C.c_uint32), ("N", C.c_uint32), ("nnz", C.c_uint32),
                ("num_row_panels", C.c_uint32), ("num_clusters", C.c_int32), ("num_clusters_true", C.c_int32),
                ("block_size", C.c_uint32), ("num_dense_blocks", C.c_uint32), ("num_dense_tiles", C.c_uint32),
                ("num_dense_values", C.c_uint64), ("num_sparse_values", C.c_uint64),
                ("row_reordering_ms", C.c_float), ("col_reordering_ms", C.c_float), ("format_build_ms", C.c_float),
                ("cluster_kernel_ms", C.c_float),
                ("num_row_groups", C.c_uint32), ("num_wide_groups", C.c_uint32), ("num_wide_tiles", C.c_uint32),
                ("num_block_tiles", C.c_uint32), ("num_wide_values", C.c_uint64), ("num_block_values", C.c_uint64),
                ("num_residual_values", C.c_uint64), ("wide_format_ms", C.c_float)]


class ShardTimes(C.Structure):
    _fields_ = [("h2d_a_ms", C.c_float), ("h2d_b_ms", C.c_float), ("allgather_b_ms", C.c_float), ("kernel_ms", C.c_float),
                ("pack_ms", C.c_float), ("gather_p_ms", C.c_float), ("unpermute_ms", C.c_float), ("d2h_ms", C.c_float),
                ("total_ms", C.c_float), ("shard_nnz", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("allgather_b_bytes", C.c_uint64), ("gather_p_bytes", C.c_uint64)]


class ReorderStats(C.Structure):
    _fields_ = [("num_dense_blocks", C.c_int32), ("average_density", C.c_float),
                ("num_dense_thread_blocks", C.c_int32), ("num_sparse_thread_blocks", C.c_int32),
                ("num_dense_data", C.c_int32), ("num_sparse_data", C.c_int32),
                ("original_num_dense_blocks", C.c_int32), ("original_average_density", C.c_float)]


def build_library():
    """Compile libbsmr_b200.so in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(HERE, "csrc")])


_lib = None


def lib():
    """The loaded C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        # BSMR_B200_LIB: another build of the same library (the -DBSMR_DEBUG build with the probe hooks, make DEBUG=1)
        path = os.environ.get("BSMR_B200_LIB", LIB_PATH)
        if not os.path.exists(path):
            raise FileNotFoundError(
                path + " is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(path)
        if path != LIB_PATH:
            # a probe build may predate part of the ABI (A/B runs against an older library): missing entry points raise
            # when they are called, not when the library is loaded
            class _Tolerant:
                def __init__(self, lib):
                    object.__setattr__(self, "_lib", lib)

                def __getattr__(self, name):
                    try:
                        return getattr(self._lib, name)
                    except AttributeError:
                        def missing(*a, **k):
                            raise AttributeError(path + " does not export " + name)
                        missing.argtypes = missing.restype = None
                        object.__setattr__(self, name, missing)
                        return missing
            L = _Tolerant(L)
        vp, cp = C.c_void_p, C.c_char_p
        L.bsmr_version.restype = cp
        L.bsmr_last_error.restype = cp
        L.bsmr_status_string.restype = cp
        L.bsmr_status_string.argtypes = [C.c_int]
        L.bsmr_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
        L.bsmr_ctx_destroy.argtypes = [vp]
        L.bsmr_ctx_synchronize.argtypes = [vp]
        L.bsmr_ctx_device_name.argtypes = [vp, C.c_char_p, C.c_size_t]
        L.bsmr_ctx_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.bsmr_calculate_block_size.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint64, u32p]
        L.bsmr_plan_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_int, C.POINTER(vp)]
        L.bsmr_plan_destroy.argtypes = [vp]
        L.bsmr_plan_row_reorder.argtypes = [vp, C.c_float, C.c_uint32, C.c_uint32]
        L.bsmr_plan_set_row_order.argtypes = [vp, u32p, C.c_uint32]
        L.bsmr_plan_col_reorder.argtypes = [vp, C.c_float]
        L.bsmr_plan_reorder.argtypes = [vp, C.c_float, C.c_float, C.c_uint32, C.c_uint32]
        L.bsmr_plan_vector_size.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
        L.bsmr_plan_vector_copy.argtypes = [vp, C.c_int, u32p, C.c_uint64]
        L.bsmr_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
        L.bsmr_plan_set_shard.argtypes = [vp, C.c_uint32, C.c_uint32, u32p, u32p, C.POINTER(C.c_uint64)]
        L.bsmr_sddmm.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_int, C.c_uint32, f32p]
        L.bsmr_sddmm_host.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_int, C.c_uint32, f32p, f32p]
        L.bsmr_sddmm_host_submit.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.POINTER(C.c_uint64)]
        L.bsmr_sddmm_host_wait.argtypes = [vp, C.c_uint64]
        L.bsmr_sddmm_batch.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint32, f32p]
        L.bsmr_sddmm_host_batch.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint32, f32p]
        L.bsmr_plan_fingerprint.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.bsmr_plan_execution_choice.argtypes = [vp, C.c_uint32, u32p]
        L.bsmr_plan_save_row_order.argtypes = [vp, C.c_char_p, C.c_float, C.c_uint32]
        L.bsmr_plan_load_row_order.argtypes = [vp, C.c_char_p, C.c_float, C.c_uint32]
        L.bsmr_sddmm_profile.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_uint32, f32p, f32p]
        L.bsmr_sddmm_profile3.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_uint32, f32p, f32p, f32p]
        L.bsmr_plan_set_wide_ratio.argtypes = [vp, C.c_float]
        L.bsmr_plan_set_l2_policy.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32]
        L.bsmr_plan_evaluate.argtypes = [vp, C.c_float, C.POINTER(ReorderStats)]
        L.bsmr_plan_autotune.argtypes = [vp, C.c_uint32, vp, vp, vp, u32p]
        L.bsmr_plan_set_execution_choice.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.bsmr_plan_fit_tile_work.argtypes = [vp, C.c_uint32, vp, vp, vp, f32p]
        L.bsmr_plan_set_tile_work.argtypes = [vp, C.c_float]
        L.bsmr_plan_set_wide_epilogue.argtypes = [vp, C.c_int]
        L.bsmr_ctx_set_host_copy_duplex.argtypes = [vp, C.c_int]
        L.bsmr_batched_transpose.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
        L.bsmr_convert_f32_to_f16.argtypes = [vp, vp, vp, C.c_uint64]
        L.bsmr_sddmm_f16b.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_int, C.c_uint32, f32p]
        L.bsmr_comm_unique_id.argtypes = [vp]
        L.bsmr_ctx_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
        L.bsmr_ctx_comm_destroy.argtypes = [vp]
        L.bsmr_ctx_comm_bcast.argtypes = [vp, vp, C.c_uint64, C.c_int]
        L.bsmr_plan_bcast_row_order.argtypes = [vp, C.c_int]
        L.bsmr_sddmm_sharded.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.c_int, C.POINTER(ShardTimes)]
        L.bsmr_sddmm_sharded_host.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.c_int, C.POINTER(ShardTimes)]
        for name in ("bsmr_ctx_create", "bsmr_ctx_destroy", "bsmr_ctx_synchronize", "bsmr_ctx_device_name",
                     "bsmr_ctx_launch_count", "bsmr_calculate_block_size", "bsmr_plan_create", "bsmr_plan_destroy",
                     "bsmr_plan_row_reorder", "bsmr_plan_set_row_order", "bsmr_plan_col_reorder", "bsmr_plan_reorder",
                     "bsmr_plan_vector_size", "bsmr_plan_vector_copy", "bsmr_plan_get_info", "bsmr_plan_set_shard",
                     "bsmr_sddmm", "bsmr_sddmm_host", "bsmr_sddmm_profile", "bsmr_sddmm_profile3",
                     "bsmr_sddmm_host_submit", "bsmr_sddmm_host_wait", "bsmr_sddmm_batch", "bsmr_sddmm_host_batch",
                     "bsmr_plan_fingerprint", "bsmr_plan_save_row_order", "bsmr_plan_load_row_order", "bsmr_plan_execution_choice",
                     "bsmr_plan_set_wide_ratio", "bsmr_plan_set_l2_policy", "bsmr_plan_evaluate", "bsmr_plan_autotune",
                     "bsmr_plan_set_execution_choice", "bsmr_plan_fit_tile_work", "bsmr_plan_set_tile_work",
                     "bsmr_plan_set_wide_epilogue", "bsmr_ctx_set_host_copy_duplex",
                     "bsmr_batched_transpose", "bsmr_convert_f32_to_f16", "bsmr_sddmm_f16b", "bsmr_comm_unique_id",
                     "bsmr_ctx_comm_init", "bsmr_ctx_comm_destroy", "bsmr_ctx_comm_bcast", "bsmr_plan_bcast_row_order",
                     "bsmr_sddmm_sharded", "bsmr_sddmm_sharded_host"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def _check(status):
    if status != 0:
        raise BsmrError(status, lib().bsmr_last_error().decode(errors="replace"))


def _ptr(x):
    """Raw address of a numpy array, a torch tensor (host or device), or an int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


class Context:
    """One device + one stream (bsmr_ctx)."""

    def __init__(self, device=0, stream=None):
        """stream: a cudaStream_t address.  None -> torch's current stream on that device when torch is imported and has
        a GPU (work queued by torch before a call is then ordered before it, and vice versa); a stream owned by the
        context otherwise."""
        if stream is None:
            import sys
            torch = sys.modules.get("torch")
            if torch is not None and torch.cuda.is_available():
                # torch's legacy default stream has handle 0, which the C ABI reads as "own stream": name it explicitly
                stream = torch.cuda.current_stream(device).cuda_stream or CUDA_STREAM_LEGACY
        self._h = C.c_void_p()
        _check(lib().bsmr_ctx_create(device, stream, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().bsmr_ctx_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def synchronize(self):
        _check(lib().bsmr_ctx_synchronize(self._h))

    def device_name(self):
        buf = C.create_string_buffer(256)
        _check(lib().bsmr_ctx_device_name(self._h, buf, 256))
        return buf.value.decode()

    def launch_count(self):
        n = C.c_uint64()
        _check(lib().bsmr_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def calculate_block_size(self, M, N, free_mem_bytes=0):
        out = C.c_uint32()
        _check(lib().bsmr_calculate_block_size(self._h, M, N, free_mem_bytes, C.byref(out)))
        return out.value

    def batched_transpose(self, width, height, num_batches, d_in, d_out):
        """batchedMatrixTranspose: every batch element height x width row-major -> width x height."""
        _check(lib().bsmr_batched_transpose(self._h, width, height, num_batches, _ptr(d_in), _ptr(d_out)))

    def convert_f32_to_f16(self, d_src, d_dst, count):
        _check(lib().bsmr_convert_f32_to_f16(self._h, _ptr(d_src), _ptr(d_dst), count))

    # ---- multi-GPU data plane (NCCL) ----
    def comm_init(self, unique_id, rank, world):
        """unique_id: the 128 bytes rank 0 got from comm_unique_id()."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(lib().bsmr_ctx_comm_init(self._h, buf, rank, world))

    def comm_destroy(self):
        _check(lib().bsmr_ctx_comm_destroy(self._h))

    def comm_bcast(self, d_tensor, nbytes, root=0):
        _check(lib().bsmr_ctx_comm_bcast(self._h, _ptr(d_tensor), nbytes, root))


def comm_unique_id():
    """Rank 0: a fresh NCCL unique id (128 bytes) to hand to Context.comm_init on every rank."""
    buf = C.create_string_buffer(128)
    _check(lib().bsmr_comm_unique_id(buf))
    return buf.raw


class Plan:
    """One sparsity pattern: the reference's BSMR object + RPHM device format (bsmr_plan)."""

    def __init__(self, ctx, M, N, row_offsets, col_indices, on_device=False):
        self.ctx = ctx
        self.M, self.N = int(M), int(N)
        if not on_device:
            row_offsets = np.ascontiguousarray(row_offsets, dtype=np.uint32)
            col_indices = np.ascontiguousarray(col_indices, dtype=np.uint32)
            self.nnz = int(len(col_indices))
        else:
            self.nnz = int(col_indices.numel())
        self._keep = (row_offsets, col_indices)
        self._h = C.c_void_p()
        _check(lib().bsmr_plan_create(ctx._h, self.M, self.N, self.nnz, _ptr(row_offsets), _ptr(col_indices),
                                      int(on_device), C.byref(self._h)))
        self._keep = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().bsmr_plan_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # ---- BSMR ----
    def row_reorder(self, alpha, block_size=0, flags=ROW_REFERENCE_COMPAT):
        _check(lib().bsmr_plan_row_reorder(self._h, alpha, block_size, flags))

    def set_row_order(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        _check(lib().bsmr_plan_set_row_order(self._h, rows.ctypes.data_as(u32p), len(rows)))

    def set_wide_ratio(self, ratio):
        """Policy of the wide row-group path, applied at the next column reorder (<= 0 disables it)."""
        _check(lib().bsmr_plan_set_wide_ratio(self._h, ratio))

    def set_wide_epilogue(self, form):
        """0 = by the fill of the wide tiles, 1 = per-entry lists, 2 = row masks; applied at the next column reorder."""
        _check(lib().bsmr_plan_set_wide_epilogue(self._h, form))

    def set_l2_policy(self, hot_budget_mb=64, min_b_mb=2048, cold_first=True):
        """Residual kernel: keep the hub columns' K-vectors (up to hot_budget_mb MiB) in L2 when B exceeds min_b_mb MiB."""
        _check(lib().bsmr_plan_set_l2_policy(self._h, hot_budget_mb, min_b_mb, int(cold_first)))

    def col_reorder(self, delta):
        _check(lib().bsmr_plan_col_reorder(self._h, delta))

    def reorder(self, alpha, delta, block_size=0, flags=ROW_REFERENCE_COMPAT):
        _check(lib().bsmr_plan_reorder(self._h, alpha, delta, block_size, flags))

    def vector(self, name):
        which = VEC[name]
        n = C.c_uint64()
        _check(lib().bsmr_plan_vector_size(self._h, which, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        _check(lib().bsmr_plan_vector_copy(self._h, which, out.ctypes.data_as(u32p), n.value))
        return out

    def info(self):
        info = PlanInfo()
        _check(lib().bsmr_plan_get_info(self._h, C.byref(info)))
        return {k: getattr(info, k) for k, _ in PlanInfo._fields_}

    def evaluate(self, delta):
        st = ReorderStats()
        _check(lib().bsmr_plan_evaluate(self._h, delta, C.byref(st)))
        return {k: getattr(st, k) for k, _ in ReorderStats._fields_}

    def set_shard(self, rank, world):
        a, b, n = C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(lib().bsmr_plan_set_shard(self._h, rank, world, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    # ---- SDDMM ----
    def sddmm(self, K, dA, dB, dP, iterations=1, flags=SDDMM_DEFAULT, timed=True):
        """Device pointers (torch CUDA tensors or raw addresses); returns ms per iteration."""
        ms = C.c_float(0)
        _check(lib().bsmr_sddmm(self._h, K, _ptr(dA), _ptr(dB), _ptr(dP), iterations, flags,
                                C.byref(ms) if timed else None))
        return ms.value

    def sddmm_profile(self, K, dA, dB, dP, flags=SDDMM_DEFAULT):
        """One pass, the two kernels timed separately: (dense_ms, residual_ms)."""
        a, b = C.c_float(0), C.c_float(0)
        _check(lib().bsmr_sddmm_profile(self._h, K, _ptr(dA), _ptr(dB), _ptr(dP), flags, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sddmm_profile3(self, K, dA, dB, dP, flags=SDDMM_DEFAULT):
        """One pass, the three kernels timed separately: (wide_ms, dense_block_ms, residual_ms)."""
        w, a, b = C.c_float(0), C.c_float(0), C.c_float(0)
        _check(lib().bsmr_sddmm_profile3(self._h, K, _ptr(dA), _ptr(dB), _ptr(dP), flags, C.byref(w), C.byref(a), C.byref(b)))
        return w.value, a.value, b.value

    def sddmm_host(self, K, hA, hB, hP=None, iterations=1, flags=SDDMM_DEFAULT):
        """Host buffers (numpy or pinned torch tensors); H2D/D2H inside.  Returns (P, kernel_ms, total_ms)."""
        if isinstance(hA, np.ndarray):
            hA = np.ascontiguousarray(hA, dtype=np.float32)
        if isinstance(hB, np.ndarray):
            hB = np.ascontiguousarray(hB, dtype=np.float32)
        if hP is None:
            hP = np.zeros(self.nnz, dtype=np.float32)
        ms, tot = C.c_float(0), C.c_float(0)
        _check(lib().bsmr_sddmm_host(self._h, K, _ptr(hA), _ptr(hB), _ptr(hP), iterations, flags, C.byref(ms),
                                     C.byref(tot)))
        return hP, ms.value, tot.value

    def execution_choice(self, K):
        """SDDMM flags a default call runs for this K (0 = three-kernel plan, unless autotune / set_execution_choice installed another)."""
        f = C.c_uint32(0)
        _check(lib().bsmr_plan_execution_choice(self._h, K, C.byref(f)))
        return f.value

    def autotune(self, K, dA, dB, dP):
        """Explicit measurement of the execution plans for this K (synchronises, writes P); installs and returns the winner."""
        f = C.c_uint32(0)
        _check(lib().bsmr_plan_autotune(self._h, K, _ptr(dA), _ptr(dB), _ptr(dP), C.byref(f)))
        return f.value

    def set_execution_choice(self, K, flags):
        _check(lib().bsmr_plan_set_execution_choice(self._h, K, flags))

    def fit_tile_work(self, K, dA, dB, dP):
        v = C.c_float(0)
        _check(lib().bsmr_plan_fit_tile_work(self._h, K, _ptr(dA), _ptr(dB), _ptr(dP), C.byref(v)))
        return v.value

    def set_tile_work(self, v):
        _check(lib().bsmr_plan_set_tile_work(self._h, v))

    def sddmm_f16b(self, K, dA, dB_f16, dP, iterations=1, flags=SDDMM_DEFAULT, timed=True):
        """B stored as fp16 ([N, K] halves), A fp32, fp32 accumulate; every nnz through the CUDA-core kernel."""
        ms = C.c_float(0)
        _check(lib().bsmr_sddmm_f16b(self._h, K, _ptr(dA), _ptr(dB_f16), _ptr(dP), iterations, flags,
                                     C.byref(ms) if timed else None))
        return ms.value

    # ---- multi-GPU data plane ----
    def bcast_row_order(self, root=0):
        _check(lib().bsmr_plan_bcast_row_order(self._h, root))

    def sddmm_sharded(self, K, dA, dB, dP_root, flags=SDDMM_DEFAULT, root=0, timed=True):
        t = ShardTimes()
        _check(lib().bsmr_sddmm_sharded(self._h, K, _ptr(dA), _ptr(dB), _ptr(dP_root), flags, root, C.byref(t) if timed else None))
        return {k: getattr(t, k) for k, _ in ShardTimes._fields_}

    def sddmm_sharded_host(self, K, hA, hB, hP, flags=SDDMM_DEFAULT, root=0):
        t = ShardTimes()
        _check(lib().bsmr_sddmm_sharded_host(self._h, K, _ptr(hA), _ptr(hB), _ptr(hP), flags, root, C.byref(t)))
        return {k: getattr(t, k) for k, _ in ShardTimes._fields_}

    def fingerprint(self):
        """64-bit fingerprint of the sparsity pattern (key of the reorder cache)."""
        h = C.c_uint64(0)
        _check(lib().bsmr_plan_fingerprint(self._h, C.byref(h)))
        return h.value

    def save_row_order(self, path, alpha, flags=ROW_REFERENCE_COMPAT):
        _check(lib().bsmr_plan_save_row_order(self._h, os.fsencode(path), alpha, flags))

    def load_row_order(self, path, alpha, flags=ROW_REFERENCE_COMPAT):
        """Install a saved row order (same pattern, alpha and flags, else BsmrError); then call col_reorder(delta)."""
        _check(lib().bsmr_plan_load_row_order(self._h, os.fsencode(path), alpha, flags))

    def sddmm_host_submit(self, K, hA, hB, hP, flags=SDDMM_DEFAULT):
        """Pipelined host-data call (pinned host buffers): queues H2D -> kernels -> D2H and returns a ticket."""
        t = C.c_uint64(0)
        _check(lib().bsmr_sddmm_host_submit(self._h, K, _ptr(hA), _ptr(hB), _ptr(hP), flags, C.byref(t)))
        return t.value

    def sddmm_host_wait(self, ticket=TICKET_ALL):
        _check(lib().bsmr_sddmm_host_wait(self._h, ticket))

    def sddmm_batch(self, num_batch, K, dA, dB, dP, flags=SDDMM_DEFAULT, timed=True):
        """sddmm_gpu_batch: device pointers, batch b at dA + b*M*K, dB + b*N*K, dP + b*nnz.  Returns total ms."""
        ms = C.c_float(0)
        _check(lib().bsmr_sddmm_batch(self._h, num_batch, K, _ptr(dA), _ptr(dB), _ptr(dP), flags, C.byref(ms) if timed else None))
        return ms.value

    def sddmm_host_batch(self, num_batch, K, hA, hB, hP, flags=SDDMM_DEFAULT):
        """The batch with host buffers, pipelined over the batch elements.  Returns wall ms."""
        ms = C.c_float(0)
        _check(lib().bsmr_sddmm_host_batch(self._h, num_batch, K, _ptr(hA), _ptr(hB), _ptr(hP), flags, C.byref(ms)))
        return ms.value


from . import synth  # noqa: E402,F401  (seeded synthetic inputs: numpy only)
