// Row reordering (BSA clustering) on the GPU, bit-compatible with the reference.
//
// Replaces bsa_rowReordering_gpu and everything under it (src/rowReordering.cu:1027-1095):
//   kernel::calculateDispersion (:49-93), the host stable sort (:1055-1062),
//   get_permutation_gpu (:893-1007) with kernel::bsa_clustering (:325-432) and
//   calculate_similarity_norm_weighted_jaccard (:235-293), the final stable sort by
//   cluster id (:986-995) and the empty-row strip (:1081-1090).
//
// The reference keeps a DENSE rows x nb encoding matrix, runs one CTA per live cluster
// (launched from the device, chained by per-row mutexes) and spends a whole-CTA reduction over
// all nb column blocks on every (cluster, row) pair.  Here:
//   * encodings are SPARSE (row -> sorted (block, count) list) built with one radix sort +
//     run-length encode; only the current cluster representative is dense, in shared memory
//   * one persistent cooperative kernel (1 CTA of 1024 threads per SM) walks the clusters in
//     order; every step evaluates the next gridDim*32 still-unassigned rows in parallel -- one
//     warp per row -- against the current representative, the first row that joins is found
//     with an atomicMin + one grid barrier, rows before it are rejected for this cluster and
//     are appended (order preserving, no scan) to the next cluster's candidate list
//   * a similarity evaluation touches only the threads of the reference CTA that own a
//     non-zero block of the candidate; all other per-thread / per-warp partial sums are those
//     of the representative alone and are computed once per representative version
// Bit-compatibility: the reference decides "sim > alpha" on fp32 values produced by a
// specific order of operations (strided per-thread sums, xor-shuffle butterfly, then a shared
// memory tree that DROPS warps when the CTA has a non-power-of-two warp count,
// include/cudaUtil.cuh:27-45).  Every float operation below is performed in that same order on
// the same operands, so the decisions -- and therefore the permutation -- are identical.
// BSMR_ROW_EXACT_REDUCE switches the lossy tree for a complete one.
//
// Two clustering kernels produce that permutation: bsa_cluster_kernel (one CTA per live cluster; inputs with long rows, where a
// per-warp dense scratch pays) and bsa_stage_kernel (one CTA per run of 32 consecutive clusters; graph-shaped inputs: 13x at
// 2^20 rows, and the only one that finishes 2^23 rows in minutes).  row_reorder() picks by the shape of the input.
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace bsmr {
namespace {

constexpr int kThreads = 256;
constexpr int kClusterThreads = 1024;
constexpr uint32_t kInf = 0xFFFFFFFFu;

inline int grid_for(uint64_t n, int per_cta, int sm_count) {
    uint64_t g = (n + per_cta - 1) / per_cta;
    const uint64_t cap = (uint64_t)sm_count * 16;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (int)g;
}

// ---- sparse encodings ------------------------------------------------------------------------
// key = row << 32 | (col / block_size), one warp per row
__global__ void enc_keys_kernel(uint32_t M, const uint32_t* __restrict__ row_offsets, const uint32_t* __restrict__ col_indices,
                                uint32_t block_size, uint64_t* __restrict__ keys) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < M; r += stride) {
        const uint32_t b = row_offsets[r], e = row_offsets[r + 1];
        for (uint32_t k = b + lane; k < e; k += 32) keys[k] = (r << 32) | (uint64_t)(col_indices[k] / block_size);
    }
}

// enc_blk[u] = column block of run u; runs per row counted with one atomic per (warp, distinct row)
__global__ void runs_split_kernel(const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ counts, uint32_t num_runs, uint32_t bd,
                                  uint32_t kept_mask, uint32_t* __restrict__ enc_blk, uint2* __restrict__ enc_pair,
                                  uint32_t* __restrict__ runs_per_row) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n32 = ((uint64_t)num_runs + 31) & ~31ull;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n32; u += (uint64_t)gridDim.x * blockDim.x) {
        const bool valid = u < num_runs;
        const uint64_t key = valid ? ukeys[u] : ~0ull;
        const uint32_t row = (uint32_t)(key >> 32);
        if (valid) {
            const uint32_t blk = (uint32_t)(key & 0xffffffffull);
            enc_blk[u] = blk;
            // {block, count | kept << 31}: what the per-thread walk of the clustering kernel reads -- whether the reference's
            // (lossy) reduction keeps the block's warp is decided here, once, instead of with a division per visit
            enc_pair[u] = make_uint2(blk, counts[u] | (((kept_mask >> ((blk % bd) >> 5)) & 1u) << 31));
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, row);
        if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(runs_per_row + row, (uint32_t)__popc(peers));
    }
}

// dispersion (src/rowReordering.cu:78-92) and the row's (possibly lossy) sum of squares, one warp per row
__global__ void dispersion_kernel(uint32_t M, const uint32_t* __restrict__ row_offsets, const uint32_t* __restrict__ enc_ptr,
                                  const uint32_t* __restrict__ enc_blk, const uint32_t* __restrict__ counts, uint32_t block_size,
                                  uint32_t bd, uint32_t kept_mask, uint32_t* __restrict__ dispersion, uint32_t* __restrict__ row_sq,
                                  uint32_t* __restrict__ row_tot) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < M; r += stride) {
        const uint32_t nz = row_offsets[r + 1] - row_offsets[r];
        const uint32_t b = enc_ptr[r], e = enc_ptr[r + 1];
        uint32_t res = 0, sq = 0, tot = 0;
        for (uint32_t j = b + lane; j < e; j += 32) {
            const uint32_t c = counts[j];
            res += block_size - c;
            if ((kept_mask >> ((enc_blk[j] % bd) >> 5)) & 1u) {
                sq += c * c;
                tot += c;
            }
        }
        res = __reduce_add_sync(0xffffffffu, res);
        sq = __reduce_add_sync(0xffffffffu, sq);
        tot = __reduce_add_sync(0xffffffffu, tot);
        if (lane == 0) {
            dispersion[r] = nz ? res + nz * (e - b) : 0u;
            row_sq[r] = sq;
            row_tot[r] = tot;          // nnz of the row in the blocks the reference's reduction keeps
        }
    }
}

__global__ void pos_info_kernel(uint32_t M, const uint32_t* __restrict__ asc, const uint32_t* __restrict__ enc_ptr,
                                const uint32_t* __restrict__ row_sq, const uint32_t* __restrict__ row_tot, uint4* __restrict__ info) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = asc[i];
        info[i] = make_uint4(row_tot[r], enc_ptr[r], enc_ptr[r + 1], row_sq[r]);
    }
}

__global__ void iota_kernel(uint32_t* p, uint32_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

__global__ void gather_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

__global__ void count_zero_kernel(const uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ out) {
    uint32_t c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) c += (v[i] == 0);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

struct Low32Key {
    __host__ __device__ uint64_t operator()(uint64_t k) const { return k; }
};

// first position in [lo, hi) whose value is >= target; warp-cooperative 32-ary search over a sorted array,
// every lane returns the same result
__device__ __forceinline__ uint32_t warp_lower_bound(const uint32_t* __restrict__ a, uint32_t lo, uint32_t hi, uint32_t target,
                                                     uint32_t lane) {
    while (hi - lo > 32) {
        const uint32_t step = (hi - lo + 31) >> 5;
        const uint32_t pos = lo + lane * step;
        const bool less = pos < hi && a[pos] < target;
        const uint32_t k = __popc(__ballot_sync(0xffffffffu, less));   // probes are sorted: the mask is a prefix
        if (k == 0) return lo;
        const uint32_t nlo = lo + (k - 1) * step + 1;
        const uint32_t nhi = lo + k * step;
        lo = nlo;
        hi = nhi < hi ? nhi : hi;
    }
    const uint32_t pos = lo + lane;
    const bool less = pos < hi && a[pos] < target;
    return lo + __popc(__ballot_sync(0xffffffffu, less));
}

// ---- the clustering kernel ------------------------------------------------------------------
struct ClusterParams {
    uint32_t M;            // rows (positions in dispersion order)
    uint32_t nb;           // column blocks per row
    uint32_t bd;           // CTA size the REFERENCE would use (src/rowReordering.cu:911-920); defines the sum order
    uint32_t first_stride; // first stride of the shared-memory tree (bd/64 in the reference)
    uint32_t zero_rows;    // leading empty rows (cluster 0)
    uint32_t list_cap;     // entries per candidate list (= M - zero_rows)
    uint32_t num_slots;    // gridDim.x + 1 candidate lists in rotation
    uint32_t kept_mask;    // reference warps that survive the (lossy) shared-memory tree
    uint32_t scratch;      // 16-bit entries of per-warp dense scratch (0 = none; else >= nb, even)
    float alpha;
    const uint32_t* asc;       // position -> row
    const uint32_t* enc_ptr;   // row -> first run
    const uint32_t* enc_blk;   // run -> column block (ascending inside a row)
    const uint32_t* counts;    // run -> nnz in the block
    const uint2* enc_pair;     // run -> {column block, nnz | kept << 31}
    uint32_t thread_prune;     // 1: one THREAD per candidate for the cheap rejections, warps only for the survivors (see the step loop)
    const uint32_t* row_sq;    // row -> (lossy) sum of squares
    const uint4* pos_info;     // position -> {nnz in kept blocks, first run, end run, (lossy) sum of squares}
    uint32_t bd_mask;          // bd - 1 when bd is a power of two (block % bd without a division), else 0
    uint32_t* repd;            // stage kernel: per CTA 32 dense representatives of nb u32 counts
    uint32_t* cluster_ids;     // position -> cluster id (pre-set: 0 for empty rows, NULL otherwise)
    uint32_t* lists;           // num_slots x list_cap positions; list of cluster c lives in slot c % num_slots
    unsigned long long* ctrl;  // per slot: list id << 33 | entries << 1 | producer-done   (single writer, release-published)
    uint32_t* status;          // [0] = number of clusters (set once), [1] = finished flag, [2] = abort (watchdog)
    unsigned long long* trace_ts; // optional: per cluster {start, first publish, end} in ns (globaltimer)
    unsigned long long* trace_stage; // optional (stage kernel): 8 counters per stage, stages <= 16384
    unsigned long long* trace; // [0] steps [1] candidates [2] joins [3] polls [4] poll cycles [5] eval cycles [6] update cycles [7] busy cycles
                               // [8] candidates rejected by the size bound alone (no block list read)
};

// block reduction with the reference's structure, executed by threads [0, bd) of the CTA
// (include/cudaUtil.cuh:13-45).  `shm` has >= 32 entries.  All kClusterThreads threads call it.
template <typename T>
__device__ __forceinline__ T ref_block_reduce(T value, T* shm, uint32_t bd, uint32_t first_stride) {
    const uint32_t tid = threadIdx.x;
    T tmp = value;
#pragma unroll
    for (int w = 1; w < 32; w <<= 1) tmp += __shfl_xor_sync(0xffffffffu, tmp, w);
    const uint32_t warp = tid >> 5, lane = tid & 31;
    if (tid < 32) shm[tid] = T(0);
    __syncthreads();
    if (lane == 0 && tid < bd) shm[warp] = tmp;
    __syncthreads();
    for (uint32_t stride = first_stride; stride >= 1; stride >>= 1) {
        if (warp < stride && lane == 0 && warp + stride < 32) shm[warp] += shm[warp + stride];
        __syncthreads();
    }
    const T r = shm[0];
    __syncthreads();
    return r;
}

// Control word of a candidate list.
__device__ __forceinline__ unsigned long long make_ctrl(uint32_t id, uint32_t count, uint32_t done) {
    return ((unsigned long long)id << 33) | ((unsigned long long)count << 1) | done;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// BSA clustering as a PIPELINE of clusters, one CTA (32 warps) per live cluster.
//
// The reference chains its cluster CTAs with one mutex per row (hand-over-hand, src/rowReordering.cu:347-357,
// 427-430): cluster c+1 may look at a row only after cluster c has passed it.  Equivalently, cluster c+1 consumes
// -- in order -- exactly the rows cluster c rejected, and its representative is the first of them.  Here that stream
// is explicit: cluster c appends the rows it rejects to list c+1 and publishes the length with a release store;
// the CTA that owns cluster c+1 polls the control word and evaluates whatever has arrived.  CTA b owns clusters
// b+1, b+1+G, b+1+2G, ... (G = gridDim.x, all CTAs co-resident: cooperative launch), so a cluster's parent is always
// owned by a CTA that is running or done -- the chain cannot deadlock -- and G+1 list buffers suffice.
// Inside a CTA every step evaluates up to 32*cpw of the arrived candidates in parallel (one warp per candidate,
// cpw candidates per warp) against the current representative; the first that joins is absorbed, the ones before it
// are rejected (appended to the child's list), the ones after it are re-evaluated against the new representative.
// cpw doubles after a step without a join and falls back to 1 after a join.
// A candidate that shares no column block with the representative has min-sum 0, i.e. similarity 0 <= alpha, and is
// rejected after one pass over its block list (alpha >= 0 only).
__global__ void __launch_bounds__(kClusterThreads, 1) bsa_cluster_kernel(ClusterParams p) {
    extern __shared__ uint32_t smem[];
    uint32_t* rep = smem;                                        // [nb]  representative encoding
    float* repn = reinterpret_cast<float*>(smem + p.nb);         // [nb]  (float)rep / norm_rep
    float* part_max = repn + p.nb;                               // [1024] per-reference-thread max partial of the representative alone
    float* warp_max = part_max + 1024;                           // [32]  the same after the warp butterfly
    // per evaluating warp: the candidate's encoding expanded to nb 16-bit counts (only when p.scratch != 0)
    uint16_t* scratch = reinterpret_cast<uint16_t*>(warp_max + 32) + (size_t)(threadIdx.x >> 5) * p.scratch;
    __shared__ uint32_t s_sq_rep, s_tot_rep;
    __shared__ float s_l1_rep;             // sum of the representative's normalised kept entries = s_tot_rep / sqrt(s_sq_rep)
    __shared__ unsigned long long s_ctrl;
    __shared__ uint32_t s_joined[2][8];    // bit k: candidate k of the step joins (up to 256 candidates); double buffered
    __shared__ uint32_t s_inscratch[2][8]; // bit k: ... and its encoding is expanded in the evaluating warp's scratch
    __shared__ uint32_t s_stop;
    __shared__ uint32_t s_maybe[2][32];    // thread-prune steps: bit l of word w = candidate 32 w + l survived the cheap tests; double buffered
    __shared__ uint32_t s_first_join[2];   // ... first candidate of the step that joins (atomicMin), double buffered
    __shared__ uint32_t s_cpw;             // candidates per warp of the next step (thread 0 decides from the step's duration)

    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nw = p.bd >> 5;
    const bool prune = p.alpha >= 0.0f;
    // integer form of the cheap upper bound  sim <= (nnz of the row in blocks shared with the representative) / (nnz of
    // the row), both over the blocks the reference's reduction keeps: reject when shared * 2^20 < (alpha - 1e-3) * total * 2^20
    const float bound = p.alpha - 1e-3f;
    constexpr uint32_t kMaxCpw = 8;
    constexpr uint32_t kWarps = kClusterThreads / 32;

    // everything derived from `rep`; called by the whole CTA after rep changed (rep complete and visible on entry)
    auto refresh = [&]() {
        if (tid == 0) { s_sq_rep = 0; s_tot_rep = 0; }
        __syncthreads();
        // sum of squares: thread t of the reference CTA owns blocks t, t+bd, ...; int e*e wraps (:241-249).  Integer
        // sums are order independent, so the reference's tree (which keeps each warp at most once) is the sum of the
        // warp sums of the kept warps.
        uint32_t sq = 0, tt = 0;
        if (tid < p.bd)
            for (uint32_t i = tid; i < p.nb; i += p.bd) { sq += rep[i] * rep[i]; tt += rep[i]; }
        sq = __reduce_add_sync(0xffffffffu, sq);
        tt = __reduce_add_sync(0xffffffffu, tt);
        if (lane == 0 && tid < p.bd && ((p.kept_mask >> wid) & 1u)) { atomicAdd(&s_sq_rep, sq); atomicAdd(&s_tot_rep, tt); }
        __syncthreads();
        const float nr = sqrtf((float)s_sq_rep);
        if (tid == 0) s_l1_rep = s_sq_rep ? (float)s_tot_rep / nr : 0.f;
        // normalised representative and, in the same pass, thread t's max partial of the representative alone (:273-282)
        float mx = 0.f;
        if (tid < p.bd) {
            for (uint32_t i = tid; i < p.nb; i += p.bd) {
                const float v = (float)rep[i] / nr;
                repn[i] = v;
                mx += v;
            }
            part_max[tid] = mx;
        }
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) mx += __shfl_xor_sync(0xffffffffu, mx, w);
        if (lane == 0) warp_max[wid] = tid < p.bd ? mx : 0.f;
        __syncthreads();
    };
    auto absorb = [&](uint32_t b, uint32_t e, bool assign) {
        if (assign) {
            for (uint32_t i = tid; i < p.nb; i += kClusterThreads) rep[i] = 0;
            __syncthreads();
        }
        for (uint32_t j = b + tid; j < e; j += kClusterThreads) rep[p.enc_blk[j]] += p.counts[j];
    };
    // does this candidate join the current representative?  (one warp; every lane returns the same answer)
    // Scratch cache: the candidate with list index x is always evaluated by warp x % 32, so when a join shifts the window
    // by a few rows the same warp meets the same candidate again and finds its encoding still expanded in its scratch
    // (c_idx / c_touched, warp-uniform registers): no global loads, no refill -- this is what bounds the cost of a long
    // run of consecutive joins.
    uint32_t c_idx = 0xFFFFFFFFu, c_touched = 0, c_pos = 0;
    unsigned long long size_rejects = 0;
    auto mod_bd = [&](uint32_t blk) -> uint32_t { return p.bd_mask ? (blk & p.bd_mask) : blk % p.bd; };
    uint4 c_info = make_uint4(0, 0, 0, 0);
    auto evaluate = [&](const uint4 info, const uint32_t x) -> bool {
        const uint32_t b = info.y, e = info.z, s_cmp = info.w;
        const uint32_t s_rep = s_sq_rep;
        if (s_rep == 0 && s_cmp == 0) return 1.0f > p.alpha;       // :258-260
        if (s_rep == 0 || s_cmp == 0) return 0.0f > p.alpha;       // :261-263
        const uint32_t n = e - b;
        const float nc = sqrtf((float)s_cmp);
        if (prune) {
            // size bound: min-sum <= the smaller of the two L1 norms of the normalised (kept) encodings, max-sum >= the larger,
            // so sim <= smaller / larger.  Both norms are known without reading the row's block list: rows that are much
            // longer or much shorter than the representative (most of a power-law graph, for most clusters) go in O(1).
            const float lc = (float)info.x / nc, lr = s_l1_rep;
            if (fminf(lc, lr) < bound * fmaxf(lc, lr)) { ++size_rejects; return false; }
        }
        float my_min = 0.f, my_max = lane < nw ? warp_max[lane] : 0.f;  // lane w = reference warp w
        if (n <= 32) {
            // ---- short block list: patch only the reference threads that own one of the row's blocks ----
            const bool valid = lane < n;
            const uint32_t blk = valid ? p.enc_blk[b + lane] : 0u;
            const uint32_t cnt = valid ? p.counts[b + lane] : 0u;
            const uint32_t t = mod_bd(blk);
            const bool kept = valid && ((p.kept_mask >> (t >> 5)) & 1u);
            if (prune) {
                const uint32_t sh = __reduce_add_sync(0xffffffffu, (kept && rep[blk] != 0) ? cnt : 0u);
                const uint32_t tot = __reduce_add_sync(0xffffffffu, kept ? cnt : 0u);
                if (sh == 0 || (float)sh < bound * (float)tot) return false;
            }
            // entries of the same reference thread (blk = t, t+bd, ...) form a group; its first lane sums the thread's
            // terms in ascending block order, taking the candidate's counts from the group members (ascending lanes)
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? t : 0xFFFF0000u + lane);
            const bool leader = valid && lane == (uint32_t)(__ffs(peers) - 1);
            uint32_t rest = peers;
            float pmin = 0.f, pmax = 0.f;
            for (uint32_t m = 0; m * p.bd < p.nb; ++m) {
                const uint32_t i = t + m * p.bd;
                const int q = rest ? __ffs(rest) - 1 : 0;
                const uint32_t blk_q = __shfl_sync(0xffffffffu, blk, q);
                const uint32_t cnt_q = __shfl_sync(0xffffffffu, cnt, q);
                if (leader && i < p.nb) {
                    const float a = repn[i];
                    if (rest && blk_q == i) {
                        const float c = (float)cnt_q / nc;
                        pmin += fminf(a, c);
                        pmax += fmaxf(a, c);
                        rest &= rest - 1;
                    } else {
                        pmax += a;
                    }
                }
            }
            uint32_t touched = __reduce_or_sync(0xffffffffu, leader ? 1u << (t >> 5) : 0u);
            while (touched) {
                const uint32_t w = __ffs(touched) - 1;
                touched &= touched - 1;
                float leaf_min = 0.f, leaf_max = part_max[(w << 5) + lane];
                uint32_t sel = __ballot_sync(0xffffffffu, leader && (t >> 5) == w);
                while (sel) {
                    const int src = __ffs(sel) - 1;
                    sel &= sel - 1;
                    const uint32_t tl = __shfl_sync(0xffffffffu, t & 31u, src);
                    const float vmin = __shfl_sync(0xffffffffu, pmin, src);
                    const float vmax = __shfl_sync(0xffffffffu, pmax, src);
                    if (lane == tl) {
                        leaf_min = vmin;
                        leaf_max = vmax;
                    }
                }
#pragma unroll
                for (int x = 1; x < 32; x <<= 1) {
                    leaf_min += __shfl_xor_sync(0xffffffffu, leaf_min, x);
                    leaf_max += __shfl_xor_sync(0xffffffffu, leaf_max, x);
                }
                if (lane == w) {
                    my_min = leaf_min;
                    my_max = leaf_max;
                }
            }
        } else {
            // ---- long block list: every touched reference warp is recomputed term by term ----
            uint32_t touched = 0, sh = 0, tot = 0;
            const bool cached = p.scratch && c_idx == x;
            if (cached) {
                touched = c_touched;
            } else {
            if (p.scratch) {
                // zero the warp's dense scratch while the first entries are in flight
                for (uint32_t i = lane; i < p.scratch / 2; i += 32) reinterpret_cast<uint32_t*>(scratch)[i] = 0u;
                __syncwarp();
            }
            for (uint32_t j0 = b + lane; j0 < e; j0 += 128) {
                uint32_t blk4[4], cnt4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {                // four independent loads in flight per array
                    const uint32_t j = j0 + 32 * u;
                    blk4[u] = j < e ? __ldg(p.enc_blk + j) : 0xFFFFFFFFu;
                    cnt4[u] = j < e ? __ldg(p.counts + j) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (blk4[u] == 0xFFFFFFFFu) continue;
                    const uint32_t w = mod_bd(blk4[u]) >> 5;
                    touched |= 1u << w;
                    if ((p.kept_mask >> w) & 1u) {
                        tot += cnt4[u];
                        if (rep[blk4[u]] != 0) sh += cnt4[u];
                    }
                    if (p.scratch) scratch[blk4[u]] = (uint16_t)cnt4[u];
                }
            }
            touched = __reduce_or_sync(0xffffffffu, touched);
            if (p.scratch) {
                c_idx = x;
                c_touched = touched;
            }
            if (prune) {
                sh = __reduce_add_sync(0xffffffffu, sh);
                tot = __reduce_add_sync(0xffffffffu, tot);
                if (sh == 0 || (float)sh < bound * (float)tot) return false;
            }
            }
            if (p.scratch) {
                // the row was expanded into the warp's dense scratch above: every term of every reference thread is one
                // shared-memory read (no search); counts fit 16 bits because a count never exceeds the block size
                __syncwarp();
                while (touched) {
                    const uint32_t w = __ffs(touched) - 1;
                    touched &= touched - 1;
                    float acc_min = 0.f, acc_max = 0.f;
                    for (uint32_t i = (w << 5) + lane; i < p.nb; i += p.bd) {   // reference thread t = w*32 + lane
                        const float a = repn[i];
                        const uint32_t cnt = scratch[i];
                        if (cnt) {
                            const float c = (float)cnt / nc;
                            acc_min += fminf(a, c);
                            acc_max += fmaxf(a, c);
                        } else {
                            acc_max += a;
                        }
                    }
#pragma unroll
                    for (int x = 1; x < 32; x <<= 1) {
                        acc_min += __shfl_xor_sync(0xffffffffu, acc_min, x);
                        acc_max += __shfl_xor_sync(0xffffffffu, acc_max, x);
                    }
                    if (lane == w) {
                        my_min = acc_min;
                        my_max = acc_max;
                    }
                }
                __syncwarp();
            } else
            while (touched) {
                const uint32_t w = __ffs(touched) - 1;
                touched &= touched - 1;
                // this lane emulates reference thread t = w*32 + lane: terms i = t, t+bd, ... ascending.  The 32 threads
                // of the warp own 32 CONSECUTIVE blocks per term, i.e. one contiguous slice of the row's sorted block
                // list: one cooperative search + one coalesced load per term.
                float acc_min = 0.f, acc_max = 0.f;
                uint32_t lo = b;
                for (uint32_t base = w << 5; base < p.nb; base += p.bd) {
                    lo = warp_lower_bound(p.enc_blk, lo, e, base, lane);
                    const uint32_t x = lo + lane;
                    const uint32_t blk_x = x < e ? p.enc_blk[x] : 0xFFFFFFFFu;
                    const bool in = blk_x < base + 32;
                    const uint32_t cnt_x = in ? p.counts[x] : 0u;
                    const uint32_t has = __reduce_or_sync(0xffffffffu, in ? 1u << (blk_x - base) : 0u);
                    const uint32_t cnt = __shfl_sync(0xffffffffu, cnt_x, __popc(has & ((1u << lane) - 1u)));
                    const uint32_t i = base + lane;
                    if (i < p.nb) {
                        const float a = repn[i];
                        if ((has >> lane) & 1u) {
                            const float c = (float)cnt / nc;
                            acc_min += fminf(a, c);
                            acc_max += fmaxf(a, c);
                        } else {
                            acc_max += a;
                        }
                    }
                }
#pragma unroll
                for (int x = 1; x < 32; x <<= 1) {
                    acc_min += __shfl_xor_sync(0xffffffffu, acc_min, x);
                    acc_max += __shfl_xor_sync(0xffffffffu, acc_max, x);
                }
                if (lane == w) {
                    my_min = acc_min;
                    my_max = acc_max;
                }
            }
        }
        // shared-memory tree of the reference over the per-warp values (lossy for odd warp counts)
        for (uint32_t stride = p.first_stride; stride >= 1; stride >>= 1) {
            const float tmin = __shfl_down_sync(0xffffffffu, my_min, stride);
            const float tmax = __shfl_down_sync(0xffffffffu, my_max, stride);
            if (lane < stride && lane + stride < 32) {
                my_min += tmin;
                my_max += tmax;
            }
        }
        const float sim = __shfl_sync(0xffffffffu, my_min, 0) / __shfl_sync(0xffffffffu, my_max, 0);
        return sim > p.alpha;
    };
    // Cheap part of a similarity decision, for ONE candidate per THREAD: the zero cases, the size bound, and -- for rows of at
    // most kWalkMax blocks -- the bound  sim <= (nnz in blocks shared with the representative) / (nnz in kept blocks),
    // evaluated by walking the row's runs serially.  Same integers and the same float comparison as the warp path's prune,
    // so the same candidates are rejected; false = rejected for sure, true = a warp has to evaluate it in full.
    constexpr uint32_t kWalkMax = 64;
    auto cheap_maybe = [&](const uint4 info) -> bool {
        if (!prune) return true;
        const uint32_t s_rep = s_sq_rep, s_cmp = info.w;
        if (s_rep == 0 || s_cmp == 0) return s_rep == 0 && s_cmp == 0;    // one zero: sim = 0 <= alpha; both: sim = 1, the warp path decides
        const float lc = (float)info.x / sqrtf((float)s_cmp), lr = s_l1_rep;
        if (fminf(lc, lr) < bound * fmaxf(lc, lr)) return false;
        if (info.z - info.y > kWalkMax) return true;
        uint32_t sh = 0;
        for (uint32_t j = info.y; j < info.z; ++j) {
            const uint2 pr = __ldg(p.enc_pair + j);
            if ((pr.y >> 31) && rep[pr.x] != 0) sh += pr.y & 0x7FFFFFFFu;
        }
        return !(sh == 0 || (float)sh < bound * (float)info.x);
    };
    unsigned long long s_ctrl_copy = 0;
    // thread 0 polls the control word of list `id` until it belongs to that list and (has > have entries or is done);
    // returns false when the run is over (no such list will ever exist) or the watchdog fired
    auto poll = [&](uint32_t id, uint32_t have) -> bool {
        if (tid == 0) {
            const unsigned long long* cw = p.ctrl + (id % p.num_slots);
            uint32_t spins = 0;
            unsigned long long v, t0 = 0;
            uint32_t stop = 0;
            for (;;) {
                v = ld_acquire_u64(cw);
                if ((uint32_t)(v >> 33) == id && ((uint32_t)((v >> 1) & 0xFFFFFFFFu) > have || (v & 1ull))) break;
                if (((volatile uint32_t*)p.status)[1] | ((volatile uint32_t*)p.status)[2]) { stop = 1; break; }
                if ((++spins & 1023u) == 0) {                 // watchdog: a stalled pipeline must end as an error, not hang
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    if (now - t0 > 120ull * 1000000000ull) { atomicExch(p.status + 2, 1u); stop = 1; break; }
                }
                __nanosleep(spins < 64 ? 20 : 200);
            }
            s_ctrl = v;
            s_stop = stop;
        }
        __syncthreads();
        const bool ok = s_stop == 0;
        const unsigned long long v = s_ctrl;
        __syncthreads();
        s_ctrl_copy = v;
        return ok;
    };

    if (tid < 16) {
        s_joined[tid >> 3][tid & 7] = 0;
        s_inscratch[tid >> 3][tid & 7] = 0;
    }
    if (tid < 2) s_first_join[tid] = 0xFFFFFFFFu;
    uint32_t par_a = 0, par_b = 0;
    uint32_t parity = 0;
    unsigned long long tr_steps = 0, tr_cand = 0, tr_joins = 0, tr_polls = 0, tr_poll_cyc = 0, tr_eval_cyc = 0, tr_upd_cyc = 0;
    const long long tr_begin = clock64();
    auto now_ns = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    auto flush_trace = [&]() {
        if (tid == 0 && p.trace) {
            atomicAdd(p.trace + 0, tr_steps); atomicAdd(p.trace + 1, tr_cand); atomicAdd(p.trace + 2, tr_joins);
            atomicAdd(p.trace + 3, tr_polls); atomicAdd(p.trace + 4, tr_poll_cyc); atomicAdd(p.trace + 5, tr_eval_cyc);
            atomicAdd(p.trace + 6, tr_upd_cyc); atomicAdd(p.trace + 7, (unsigned long long)(clock64() - tr_begin));
            atomicAdd(p.trace + 8, size_rejects);
        }
    };
    for (uint32_t c = blockIdx.x + 1;; c += gridDim.x) {
        const uint32_t* in = p.lists + (size_t)(c % p.num_slots) * p.list_cap;
        uint32_t* out = p.lists + (size_t)((c + 1) % p.num_slots) * p.list_cap;
        unsigned long long* out_ctrl = p.ctrl + ((c + 1) % p.num_slots);
        {
            const long long t0 = clock64();
            const bool ok = poll(c, 0);
            tr_poll_cyc += clock64() - t0;
            ++tr_polls;
            if (!ok) { flush_trace(); return; }
        }
        uint32_t avail = (uint32_t)((s_ctrl_copy >> 1) & 0xFFFFFFFFu);
        bool in_done = (s_ctrl_copy & 1ull) != 0;
        if (avail == 0) {
            // the parent rejected nothing: cluster c does not exist and the run is over
            if (tid == 0) {
                p.status[0] = c - 1;
                __threadfence();
                atomicExch(p.status + 1, 1u);
            }
            flush_trace();
            return;
        }
        if (tid == 0) st_release_u64(out_ctrl, make_ctrl(c + 1, 0, 0));
        if (tid == 0 && p.trace_ts && c <= p.M) { p.trace_ts[3 * c] = now_ns(); p.trace_ts[3 * c + 1] = 0; }
        const uint32_t start_pos = __ldcg(in);
        if (tid == 0) p.cluster_ids[start_pos] = c;
        {
            const uint4 si = __ldg(p.pos_info + start_pos);
            absorb(si.y, si.z, true);
            __syncthreads();
            refresh();
        }
        uint32_t cursor = 1, produced = 0, cpw = 1, published = 0;
        c_idx = 0xFFFFFFFFu;                                   // list indices of the previous cluster mean nothing here
        for (;;) {
            if (cursor >= avail) {
                if (in_done) break;
                {
                    const long long t0 = clock64();
                    const bool ok = poll(c, cursor);
                    tr_poll_cyc += clock64() - t0;
                    ++tr_polls;
                    if (!ok) { flush_trace(); return; }
                }
                avail = (uint32_t)((s_ctrl_copy >> 1) & 0xFFFFFFFFu);
                in_done = (s_ctrl_copy & 1ull) != 0;
                continue;
            }
            if (p.thread_prune) {
                // ---- thread-prune step: up to 1024 candidates, one per thread ------------------------------------------------
                // Measured on R-MAT graphs (ncu, profiles/r02d_*): the warp-per-candidate step spends ~300 warp instructions
                // on a (cluster, row) pair that the bounds reject anyway -- 99.9 % of the pairs -- and the whole pipeline runs
                // at the pace of that bookkeeping.  Here a THREAD applies the bounds to its candidate (the size bound in O(1),
                // the shared-nnz bound by walking the row's <= 64 runs); warps evaluate only the survivors, in order, up to the
                // first that joins.  The rejections are the same integers and comparisons, so the permutation is unchanged.
                // (variants measured on the 2^20-row graph, profiles/r02m_*, r02n_*: 8 survivors per warp with a warp-wide prune for
                // the long rows 67.9 s, walks of up to 256 runs 45.9 s, this one 38.8 s: a warp's survivors are serial round trips, a
                // long walk idles 31 lanes, and either way a longer step starves the children)
                constexpr uint32_t kCpwB = 2;                               // survivors a warp evaluates per step
                const uint32_t take = min(avail - cursor, (uint32_t)kClusterThreads);
                ++tr_steps;
                tr_cand += take;
                uint32_t pos1 = 0xFFFFFFFFu;
                bool maybe = false;
                if (tid < take) {
                    pos1 = __ldcg(in + cursor + tid);
                    maybe = cheap_maybe(__ldg(p.pos_info + pos1));
                }
                const uint32_t mb = __ballot_sync(0xffffffffu, maybe);
                if (lane == 0) s_maybe[par_a][wid] = mb;
                __syncthreads();                                  // S1: verdicts in; the rejected rows stored by earlier steps are ordered before it
                if (tid == 0 && produced != published) st_release_u64(out_ctrl, make_ctrl(c + 1, produced, 0));
                published = produced;
                const uint32_t wmask = s_maybe[par_a][lane];
                par_a ^= 1;
                uint32_t pre = __popc(wmask);                     // inclusive prefix of the survivors over the 32 words
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, pre, d);
                    if ((int)lane >= d) pre += t;
                }
                const uint32_t total_maybe = __shfl_sync(0xffffffffu, pre, 31);
                if (total_maybe == 0) {                           // the common step: nothing survives, everything moves on to the child
                    if (tid < take) out[produced + tid] = pos1;
                    produced += take;
                    cursor += take;
                    continue;
                }
                // index (within the step) of the survivor with rank r
                auto nth_maybe = [&](uint32_t r) -> uint32_t {
                    const uint32_t w = __ffs(__ballot_sync(0xffffffffu, pre > r)) - 1;      // first word whose prefix passes r
                    const uint32_t before = __shfl_sync(0xffffffffu, pre - __popc(wmask), w);
                    const uint32_t bits = __shfl_sync(0xffffffffu, wmask, w);
                    return w * 32 + __fns(bits, 0, (int)(r - before) + 1);
                };
                const uint32_t evaluated = min(total_maybe, kWarps * kCpwB);
                const uint32_t take_eff = total_maybe > evaluated ? nth_maybe(evaluated) : take;    // later survivors wait for the next step
#pragma unroll 1
                for (uint32_t q = 0; q < kCpwB; ++q) {
                    const uint32_t r = q * kWarps + wid;
                    if (r < evaluated) {                          // warp-uniform
                        const uint32_t k = nth_maybe(r);
                        const uint32_t kpos = __ldcg(in + cursor + k);
                        if (evaluate(__ldg(p.pos_info + kpos), cursor + k) && lane == 0) atomicMin(&s_first_join[par_b], k);
                    }
                }
                __syncthreads();                                  // S2: every survivor decided
                const uint32_t fj = s_first_join[par_b];
                par_b ^= 1;
                if (tid == 0) s_first_join[par_b] = 0xFFFFFFFFu;   // last read two survivor steps ago; visible at the next S1
                const uint32_t n_rej = fj == 0xFFFFFFFFu ? take_eff : fj;
                if (tid < n_rej) out[produced + tid] = pos1;
                produced += n_rej;
                if (fj == 0xFFFFFFFFu) {
                    cursor += take_eff;
                } else {
                    if (tid == fj) p.cluster_ids[pos1] = c;
                    const uint4 ji = __ldg(p.pos_info + __ldcg(in + cursor + fj));
                    absorb(ji.y, ji.z, false);
                    refresh();                                    // its first barrier also closes the absorb
                    cursor += fj + 1;
                    ++tr_joins;
                }
                continue;
            }
            const uint32_t take = min(avail - cursor, kWarps * cpw);
            const long long tr_t0 = clock64();
            ++tr_steps;
            tr_cand += take;
            // the candidate with list index x goes to warp x % 32; fetch everything first, then evaluate
            uint32_t my_pos[kMaxCpw];
            uint4 my_info[kMaxCpw];
            const uint32_t k0 = (wid + kWarps - (cursor % kWarps)) % kWarps;     // first candidate of this step owned by this warp
            const bool hit0 = k0 < take && c_idx == cursor + k0;     // first candidate already expanded in this warp's scratch
#pragma unroll
            for (uint32_t q = 0; q < kMaxCpw; ++q) {
                const uint32_t k = q * kWarps + k0;
                my_pos[q] = (q < cpw && k < take) ? ((q == 0 && hit0) ? c_pos : __ldcg(in + cursor + k)) : 0xFFFFFFFFu;
            }
#pragma unroll
            for (uint32_t q = 0; q < kMaxCpw; ++q)
                if (my_pos[q] != 0xFFFFFFFFu) my_info[q] = (q == 0 && hit0) ? c_info : __ldg(p.pos_info + my_pos[q]);
#pragma unroll
            for (uint32_t q = 0; q < kMaxCpw; ++q) {
                if (my_pos[q] != 0xFFFFFFFFu) {
                    const uint32_t k = q * kWarps + k0;
                    const bool joins = evaluate(my_info[q], cursor + k);
                    if (c_idx == cursor + k) {                 // evaluate() left this candidate in the scratch
                        c_pos = my_pos[q];
                        c_info = my_info[q];
                    }
                    if (joins && lane == 0) {
                        atomicOr(&s_joined[parity][k >> 5], 1u << (k & 31));
                        // (with several candidates per warp a later one overwrites the scratch)
                        if (cpw == 1 && c_idx == cursor + k) atomicOr(&s_inscratch[parity][k >> 5], 1u << (k & 31));
                    }
                }
            }
            __syncthreads();                                  // #1: all verdicts in
            const long long tr_t1 = clock64();
            tr_eval_cyc += tr_t1 - tr_t0;
            if (tid == 0) {
                // a step should take a few microseconds: long steps starve the whole chain of descendants (they only
                // receive rows when this cluster publishes), short ones waste time in barriers
                const long long dur = tr_t1 - tr_t0;
                uint32_t next = cpw;
                if (dur < 6000 && cpw < kMaxCpw) next = cpw << 1;
                else if (dur > 24000 && cpw > 1) next = cpw >> 1;
                s_cpw = next;
            }
            uint32_t fj = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                const uint32_t bits = s_joined[parity][i];
                if (bits) fj = i * 32 + (__ffs(bits) - 1);
            }
            const uint32_t n_rej = fj == 0xFFFFFFFFu ? take : fj;
#pragma unroll
            for (uint32_t q = 0; q < kMaxCpw; ++q) {
                const uint32_t k = q * kWarps + k0;
                if (my_pos[q] != 0xFFFFFFFFu && k < n_rej && lane == 0) out[produced + k] = my_pos[q];
            }
            produced += n_rej;
            uint4 ji = make_uint4(0, 0, 0, 0);
            uint32_t jpos = 0;
            bool from_scratch = false;
            if (fj != 0xFFFFFFFFu) {
                from_scratch = (s_inscratch[parity][fj >> 5] >> (fj & 31)) & 1u;
                if (!from_scratch) {
                    jpos = __ldcg(in + cursor + fj);
                    ji = __ldg(p.pos_info + jpos);
                }
#pragma unroll
                for (uint32_t q = 0; q < kMaxCpw; ++q)       // the warp that evaluated the joiner records its cluster
                    if (my_pos[q] != 0xFFFFFFFFu && q * kWarps + k0 == fj && lane == 0) p.cluster_ids[my_pos[q]] = c;
            }
            __syncthreads();                                  // #2: verdicts read by everybody, rejected rows stored
            if (tid < 8) {                                    // next use of these buffers is two steps away
                s_joined[parity][tid] = 0;
                s_inscratch[parity][tid] = 0;
            }
            parity ^= 1;
            // release: the list stores of the other threads are ordered before this store by the barrier (cumulativity)
            if (tid == 0 && n_rej) {
                st_release_u64(out_ctrl, make_ctrl(c + 1, produced, 0));
                if (p.trace_ts && c <= p.M && p.trace_ts[3 * c + 1] == 0) p.trace_ts[3 * c + 1] = now_ns();
            }
            if (fj == 0xFFFFFFFFu) {
                cursor += take;
                cpw = s_cpw;
            } else {
                if (from_scratch) {
                    // the joining row is still expanded in the scratch of the warp that evaluated it
                    const uint16_t* js = reinterpret_cast<const uint16_t*>(warp_max + 32) + (size_t)((cursor + fj) % kWarps) * p.scratch;
                    for (uint32_t i = tid; i < p.nb; i += kClusterThreads) rep[i] += js[i];
                } else {
                    absorb(ji.y, ji.z, false);
                }
                refresh();                                    // its first barrier also closes the absorb
                cursor += fj + 1;
                cpw = 1;
                ++tr_joins;
                tr_upd_cyc += clock64() - tr_t1;
            }
        }
        __syncthreads();
        if (tid == 0) st_release_u64(out_ctrl, make_ctrl(c + 1, produced, 1));
        if (tid == 0 && p.trace_ts && c <= p.M) p.trace_ts[3 * c + 2] = now_ns();
    }
}

// ---- the stage kernel: one CTA per RUN of kStageReps consecutive clusters ---------------------------------------------
//
// Same pipeline between the CTAs as bsa_cluster_kernel (lists in rotation, release-published control words, cooperative
// launch), but the unit that owns a list is a STAGE: clusters (s-1)*32+1 .. s*32.  Row-major view of the reference's
// semantics: a row walks the clusters in ascending order, joins the first whose (current) representative accepts it, and
// founds the next cluster when none does.  A stage therefore keeps 32 representatives; a row is tested against all of them
// while its pos_info / block list is in the thread's registers, and is forwarded ONCE to the next stage -- the list traffic,
// the polls and the pipeline hops of 32 clusters collapse into one.
//   * the stage's representatives live as (a) `slots`: one word per column block, bit r = representative r has nnz in the
//     block (shared memory), (b) dense u32 count vectors in global memory (L2; only the full evaluation reads them),
//     (c) the per-reference-thread / per-reference-warp partial sums of the representative alone (shared memory)
//   * streaming step: up to 1024 rows, one per THREAD.  Size bound against the 32 representatives (two compares each), then
//     ONE walk over the row's runs accumulates the nnz shared with all 32 representatives at once in bit-sliced counters
//     (8 planes of 32 bits: plane q holds bit q of the 32 sums; adding a run's count under the slot word is a ripple of
//     AND / XOR), and a bit-sliced compare against the row's threshold gives the mask of representatives the row may still
//     join.  Rows with an empty mask (99.9 % on graphs) are forwarded; the others get a warp each that evaluates the
//     reference's similarity (same float operations in the same order as bsa_cluster_kernel) against the set bits in
//     ascending order.  The first row that joins ends the step: its representative is refreshed, the rows behind it are redone.
//   * founding: while the stage has fewer than 32 representatives the stream is consumed row by row: the row joins the first
//     existing representative that accepts it (the candidates are evaluated by different warps at once) or founds the next.
//   * certain joins: a representative that holds a single column block b (count c) and a row whose only run is block b (count
//     k): both normalise to exactly 1.0f in block b -- (float)c / sqrtf((float)(c * c)) == 1.0f for every c < 65536, checked
//     exhaustively -- so the similarity is 1 and, after the join, the normalised representative, its L1 norm and its partial sums
//     are bit-identical to what they were.  When the bounds leave such a row no other candidate, it joins without an
//     evaluation and WITHOUT invalidating anything computed for the rows behind it: a step absorbs any number of them.  (On
//     R-MAT graphs a third of the rows are single-run rows that pile into one cluster per hub block; handled one join per
//     step they were 56 % of the 2^20-row run.)
// All bounds are exact (they only skip pairs whose similarity cannot exceed alpha), so the permutation is the reference's.
constexpr uint32_t kStageReps = 32;
constexpr uint32_t kStageTerms = 8;    // column blocks per reference thread the stage kernel takes: nb <= 8 * bd

__global__ void __launch_bounds__(kClusterThreads, 1) bsa_stage_kernel(ClusterParams p) {
    extern __shared__ uint32_t smem[];
    const uint32_t nbp = (p.nb + 3u) & ~3u;
    uint32_t* slots = smem;                                                    // [nb]   bit r: representative r has nnz in the block
    float* part_max = reinterpret_cast<float*>(smem + nbp);                    // [32][1024] per-reference-thread max partial
    float* warp_max = part_max + kStageReps * 1024;                            // [32][32]   the same after the warp butterfly
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(warp_max + kStageReps * 32);   // [1024] this step's verdict masks
    uint16_t* scratch = reinterpret_cast<uint16_t*>(s_mask + 1024);               // [nb] one long row, expanded (zero between uses)
    __shared__ uint32_t sq_s[kStageReps];      // (lossy) sum of squares of representative r
    __shared__ uint32_t tot_s[kStageReps];     // (lossy) sum of its counts
    __shared__ uint32_t s_sh[kStageReps];      // shared_mask_scratch: nnz the row shares with representative r
    __shared__ uint32_t single_s[kStageReps];  // the one (kept) column block representative r consists of, else kNone
    __shared__ uint32_t sc_s[kStageReps];      // ... and its count there
    __shared__ uint32_t s_add[kStageReps];     // certain joins of the step: nnz added to representative r (zero between steps)
    __shared__ uint32_t s_cj[2][32];           // bit: the row of the step is a certain join
    __shared__ uint32_t s_out[2][32];          // careful mode: what the round decided for the rows of the batch
    __shared__ uint32_t s_fpos[32];            // founding: the next rows of the input, fetched together ...
    __shared__ uint4 s_finfo[32];              // ... their pos_info ...
    __shared__ uint2 s_runs64[32 * 64];        // ... and the runs of those with at most 64; streaming: warp w stages a row of 33..64 runs at [w]
    __shared__ uint32_t s_wbits[32 * 32];      // evaluate_medium, per warp: bit t = reference thread t already has its leader (zero between uses)
    __shared__ float nr_s[kStageReps];         // its square root
    __shared__ float2 l_s[kStageReps];         // {L1 norm of the normalised kept entries, bound * that}
    __shared__ uint32_t s_nz;                  // bit r: sq_s[r] != 0
    __shared__ float s_blr_min, s_lr_max;      // over the representatives with sq != 0: min of l_s.y, max of l_s.x
    __shared__ uint32_t s_red[2];
    __shared__ uint32_t s_ev[2][32];
    __shared__ uint32_t s_long[2][32];         // bit: the event row still needs the warp-level shared-nnz bound
    __shared__ uint32_t s_first[2];
    __shared__ float s_fv[32 * 32];            // found_sparse: per warp, the normalised counts of a chunk of runs
    __shared__ uint32_t s_pair[32];            // batch founding: bit i of word j = row j of the batch might join a cluster founded by row i
    __shared__ uint32_t s_anyq;                // a long row of the step is queued for the scratch pass
#ifndef BSMR_STAGE_CPW
#define BSMR_STAGE_CPW 4
#endif
    __shared__ uint8_t s_qun[32 * BSMR_STAGE_CPW];             // ... and whether the shared-nnz bound is still to be applied (very long rows)
    __shared__ uint32_t s_qmask[32 * BSMR_STAGE_CPW];           // per event row of the step: representatives left for the scratch pass (rows of > 32 runs)
    __shared__ uint32_t s_touched;             // reference warps that own a block of the row in the scratch
    __shared__ float2 s_res[kStageReps * 32];  // scratch pass: per (representative, reference warp) the {min, max} sums
    __shared__ unsigned long long s_ctrl;
    __shared__ uint32_t s_stop;

    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nw = p.bd >> 5;
    const float bound = p.alpha - 1e-3f;
    constexpr uint32_t kWarps = kClusterThreads / 32;
#ifndef BSMR_STAGE_WALK_MAX
#define BSMR_STAGE_WALK_MAX 64
#endif
    constexpr uint32_t kWalkMax = BSMR_STAGE_WALK_MAX;      // runs a thread walks itself (measured: 128 -> 3.5 s instead of 3.0 s at 2^20 rows, 25.4 instead of 23.8 s at 2^22)
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    const bool cj_on = p.alpha < 1.0f;         // similarity 1 must be an acceptance
    uint32_t* const repd = p.repd + (size_t)blockIdx.x * kStageReps * p.nb;     // dense representatives of this CTA
    auto mod_bd = [&](uint32_t blk) -> uint32_t { return p.bd_mask ? (blk & p.bd_mask) : blk % p.bd; };

    // add the row's runs to representative r (dense counts + slot bits); the caller's next barrier closes it
    auto absorb = [&](const uint4 info, uint32_t r) {
        uint32_t* rd = repd + (size_t)r * p.nb;
        for (uint32_t j = info.y + tid; j < info.z; j += kClusterThreads) {
            const uint32_t blk = __ldg(p.enc_blk + j);
            rd[blk] = __ldcg(rd + blk) + __ldg(p.counts + j);          // a row's runs are distinct blocks: no conflicts
            atomicOr(&slots[blk], 1u << r);
        }
    };
    // everything derived from representative r; called by the whole CTA after an absorb
    auto refresh = [&](uint32_t r, uint32_t nrep_now) {
        const uint32_t* rd = repd + (size_t)r * p.nb;
        if (tid == 0) { s_red[0] = 0; s_red[1] = 0; }
        __syncthreads();                                   // closes the absorb
        uint32_t sq = 0, tt = 0;
        if (tid < p.bd)
            for (uint32_t i = tid; i < p.nb; i += p.bd) { const uint32_t v = __ldcg(rd + i); sq += v * v; tt += v; }
        sq = __reduce_add_sync(0xffffffffu, sq);
        tt = __reduce_add_sync(0xffffffffu, tt);
        if (lane == 0 && tid < p.bd && ((p.kept_mask >> wid) & 1u)) { atomicAdd(&s_red[0], sq); atomicAdd(&s_red[1], tt); }
        __syncthreads();
        const uint32_t sqr = s_red[0], ttr = s_red[1];
        const float nr = sqrtf((float)sqr);
        float mx = 0.f;
        if (tid < p.bd) {
            for (uint32_t i = tid; i < p.nb; i += p.bd) mx += (float)__ldcg(rd + i) / nr;
            part_max[r * 1024 + tid] = mx;
        }
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) mx += __shfl_xor_sync(0xffffffffu, mx, w);
        if (lane == 0) warp_max[r * 32 + wid] = tid < p.bd ? mx : 0.f;
        if (tid == 0) {
            sq_s[r] = sqr;
            tot_s[r] = ttr;
            nr_s[r] = nr;
            const float l1 = sqr ? (float)ttr / nr : 0.f;
            l_s[r] = make_float2(l1, bound * l1);
            if (sqr) atomicOr(&s_nz, 1u << r);
        }
        __syncthreads();
        if (wid == 0) {
            const bool on = lane < nrep_now && ((s_nz >> lane) & 1u);
            float lo = on ? l_s[lane].y : INFINITY, hi = on ? l_s[lane].x : 0.f;
#pragma unroll
            for (int w = 1; w < 32; w <<= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, w));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, w));
            }
            if (lane == 0) { s_blr_min = lo; s_lr_max = hi; s_red[0] = 0; s_red[1] = 0; }
        }
        __syncthreads();
    };
    // Which of the first `nrep` representatives can this row still join?  Exact bounds only (a cleared bit = similarity
    // cannot exceed alpha); one thread, no synchronisation.
    // Second shared-nnz bound.  max(a, c) = a + c - min(a, c) per block, so max-sum = L1(rep) + L1(row) - min-sum and
    // sim = min-sum / (L1(rep) + L1(row) - min-sum), increasing in min-sum; min-sum <= U = the row's normalised nnz in the blocks
    // it shares with the representative.  Reject when U / (L1(rep) + L1(row) - U) < alpha - 1e-3 (all sums over the kept blocks,
    // like the reference's).  Against a representative much larger than the row this asks for nearly ALL of the row to be
    // shared, where the first bound (U / L1(row)) only asks for 30 %.
    auto keep_by_union = [&](uint32_t sh, float nc, float lc, uint32_t r) -> bool {
        const float U = (float)sh / nc;
        return !(U < bound * (l_s[r].x + lc - U));
    };
    // the zero cases and the size bound: which of the first `nrep` representatives are still possible (no block list read)
    auto size_mask = [&](const uint4 info, uint32_t nrep) -> uint32_t {
        const uint32_t live = nrep >= 32 ? 0xFFFFFFFFu : ((1u << nrep) - 1u);
        const uint32_t nz = s_nz;
        if (info.w == 0) return live & ~nz;        // zero (kept) encoding: similarity 1 with an all-zero representative, 0 otherwise (:258-263)
        const float lc = (float)info.x / sqrtf((float)info.w);
        const float blc = bound * lc;
        if (lc < s_blr_min || s_lr_max < blc) return 0u;
        uint32_t alive = 0;
        if (nrep >= 32) {
#pragma unroll
            for (uint32_t r = 0; r < 32; ++r) {
                const float2 l = l_s[r];
                alive |= (uint32_t)(lc >= l.y && l.x >= blc) << r;
            }
        } else {
            for (uint32_t r = 0; r < nrep; ++r) {
                const float2 l = l_s[r];
                alive |= (uint32_t)(lc >= l.y && l.x >= blc) << r;
            }
        }
        return alive & nz & live;
    };
    // One WARP, any row length: lane r sums the row's (kept) nnz in the blocks it shares with representative r; returns the
    // representatives of `alive` that pass  shared >= 1 and shared >= bound * (kept nnz of the row).  Not for info.w == 0.
    auto warp_shared_mask = [&](const uint4 info, uint32_t alive, const uint2* sruns) -> uint32_t {
        uint32_t sh = 0;
        for (uint32_t j0 = info.y; j0 < info.z; j0 += 128) {           // four chunks of 32 runs in flight (long rows: one round trip per 128 runs)
            uint2 pr4[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t j = j0 + 32 * u + lane;
                pr4[u] = j < info.z ? (sruns ? sruns[j - info.y] : __ldg(p.enc_pair + j)) : make_uint2(0u, 0u);   // sruns: the row's runs in shared memory
            }
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                if (j0 + 32 * u >= info.z) break;                      // (uniform)
                const uint32_t m = (pr4[u].y >> 31) ? (slots[pr4[u].x] & alive) : 0u;
                const uint32_t cnt = pr4[u].y & 0x7FFFFFFFu;
                uint32_t hit = __ballot_sync(0xffffffffu, m != 0);
                while (hit) {
                    const int i = __ffs(hit) - 1;
                    hit &= hit - 1;
                    const uint32_t mi = __shfl_sync(0xffffffffu, m, i);
                    const uint32_t ci = __shfl_sync(0xffffffffu, cnt, i);
                    if ((mi >> lane) & 1u) sh += ci;
                }
            }
        }
        const float nc = sqrtf((float)info.w);
        const bool ok = sh != 0 && !((float)sh < bound * (float)info.x) && keep_by_union(sh, nc, (float)info.x / nc, lane);
        return __ballot_sync(0xffffffffu, ok) & alive;
    };
    // One THREAD per row: size bound, then (rows of at most kWalkMax runs and fewer than 256 kept nnz) the shared-nnz bound
    // for all representatives at once in bit-sliced counters.  Longer rows come back with is_long set and the size mask: a
    // warp finishes them with warp_shared_mask.
    auto bounds_mask = [&](const uint4 info, uint32_t nrep, bool& is_long) -> uint32_t {
        is_long = false;
        const uint32_t alive = size_mask(info, nrep);
        if (alive == 0 || info.w == 0) return alive;
        if (info.z - info.y > kWalkMax || info.x >= 256u) { is_long = true; return alive; }
        uint32_t S[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) S[q] = 0;
        for (uint32_t j = info.y; j < info.z; ++j) {
            const uint2 pr = __ldg(p.enc_pair + j);
            if (pr.y >> 31) {
                const uint32_t m = slots[pr.x] & alive;
                if (m) {
                    const uint32_t cnt = pr.y & 0x7FFFFFFFu;            // < 256: the row's kept nnz are
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if ((cnt >> k) & 1u) {
                            uint32_t carry = m;
#pragma unroll
                            for (int q = k; q < 8; ++q) {
                                const uint32_t t = S[q] & carry;
                                S[q] ^= carry;
                                carry = t;
                            }
                        }
                    }
                }
            }
        }
        // survive iff sh >= 1 and (float)sh >= bound * tot, i.e. sh >= T (sums below 2^24 convert exactly)
        const float thr = bound * (float)info.x;
        uint32_t T = thr > 0.f ? (uint32_t)ceilf(thr) : 0u;
        if (T == 0) T = 1;
        if (T > 255u) return 0u;                                       // sh <= tot < 256
        uint32_t ge = 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < 8; ++q) ge = ((T >> q) & 1u) ? (S[q] & ge) : (S[q] | ge);
        uint32_t out = ge & alive;
        if (out) {                                                      // the few survivors: the second bound, per representative
            const float nc = sqrtf((float)info.w);
            const float lc = (float)info.x / nc;
            for (uint32_t rest = out; rest; rest &= rest - 1) {
                const uint32_t r = __ffs(rest) - 1;
                uint32_t sh = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) sh |= ((S[q] >> r) & 1u) << q;
                if (!keep_by_union(sh, nc, lc, r)) out &= ~(1u << r);
            }
        }
        return out;
    };
    // does the row join representative r?  One warp; the float operations of calculate_similarity_norm_weighted_jaccard
    // in the reference's order (see bsa_cluster_kernel::evaluate; here the normalised representative is recomputed from
    // the dense counts: the same division on the same operands)
    auto evaluate_pair = [&](const uint4 info, const uint32_t r, const uint2* sruns) -> bool {
        const uint32_t b = info.y, e = info.z, s_cmp = info.w;
        const uint32_t s_rep = sq_s[r];
        if (s_rep == 0 && s_cmp == 0) return 1.0f > p.alpha;
        if (s_rep == 0 || s_cmp == 0) return 0.0f > p.alpha;
        const uint32_t* rd = repd + (size_t)r * p.nb;
        const float* pm = part_max + r * 1024;
        const float nr = nr_s[r];
        const uint32_t n = e - b;
        const float nc = sqrtf((float)s_cmp);
        float my_min = 0.f, my_max = lane < nw ? warp_max[r * 32 + lane] : 0.f;
        {   // n <= 32 (longer rows: evaluate_scratch): patch only the reference threads that own one of the row's blocks
            const bool valid = lane < n;
            const uint2 spr = (valid && sruns) ? sruns[lane] : make_uint2(0u, 0u);
            const uint32_t blk = valid ? (sruns ? spr.x : __ldg(p.enc_blk + b + lane)) : 0u;
            const uint32_t cnt = valid ? (sruns ? (spr.y & 0x7FFFFFFFu) : __ldg(p.counts + b + lane)) : 0u;
            const uint32_t t = mod_bd(blk);
            const bool kept = valid && ((p.kept_mask >> (t >> 5)) & 1u);
            {
                const uint32_t sh = __reduce_add_sync(0xffffffffu, (kept && ((slots[blk] >> r) & 1u)) ? cnt : 0u);
                const uint32_t tot = __reduce_add_sync(0xffffffffu, kept ? cnt : 0u);
                if (sh == 0 || (float)sh < bound * (float)tot) return false;
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? t : 0xFFFF0000u + lane);
            const bool leader = valid && lane == (uint32_t)(__ffs(peers) - 1);
            uint32_t rest = peers;
            float pmin = 0.f, pmax = 0.f;
            for (uint32_t m = 0; m * p.bd < p.nb; ++m) {
                const uint32_t i = t + m * p.bd;
                const int q = rest ? __ffs(rest) - 1 : 0;
                const uint32_t blk_q = __shfl_sync(0xffffffffu, blk, q);
                const uint32_t cnt_q = __shfl_sync(0xffffffffu, cnt, q);
                if (leader && i < p.nb) {
                    const float a = (float)__ldcg(rd + i) / nr;
                    if (rest && blk_q == i) {
                        const float c = (float)cnt_q / nc;
                        pmin += fminf(a, c);
                        pmax += fmaxf(a, c);
                        rest &= rest - 1;
                    } else {
                        pmax += a;
                    }
                }
            }
            uint32_t touched = __reduce_or_sync(0xffffffffu, leader ? 1u << (t >> 5) : 0u);
            while (touched) {
                const uint32_t w = __ffs(touched) - 1;
                touched &= touched - 1;
                float leaf_min = 0.f, leaf_max = pm[(w << 5) + lane];
                uint32_t sel = __ballot_sync(0xffffffffu, leader && (t >> 5) == w);
                while (sel) {
                    const int src = __ffs(sel) - 1;
                    sel &= sel - 1;
                    const uint32_t tl = __shfl_sync(0xffffffffu, t & 31u, src);
                    const float vmin = __shfl_sync(0xffffffffu, pmin, src);
                    const float vmax = __shfl_sync(0xffffffffu, pmax, src);
                    if (lane == tl) {
                        leaf_min = vmin;
                        leaf_max = vmax;
                    }
                }
#pragma unroll
                for (int x = 1; x < 32; x <<= 1) {
                    leaf_min += __shfl_xor_sync(0xffffffffu, leaf_min, x);
                    leaf_max += __shfl_xor_sync(0xffffffffu, leaf_max, x);
                }
                if (lane == w) {
                    my_min = leaf_min;
                    my_max = leaf_max;
                }
            }
        }
        for (uint32_t stride = p.first_stride; stride >= 1; stride >>= 1) {
            const float tmin = __shfl_down_sync(0xffffffffu, my_min, stride);
            const float tmax = __shfl_down_sync(0xffffffffu, my_max, stride);
            if (lane < stride && lane + stride < 32) {
                my_min += tmin;
                my_max += tmax;
            }
        }
        const float sim = __shfl_sync(0xffffffffu, my_min, 0) / __shfl_sync(0xffffffffu, my_max, 0);
        return sim > p.alpha;
    };
    // Fast decision, one warp, a row of at most 64 runs staged in shared memory; lane = representative.  In real arithmetic
    // max-sum = L1(rep) + L1(row) - min-sum (see keep_by_union), and min-sum only has terms in the row's blocks: n reads of the
    // dense representative give the similarity to ~1e-6, all candidates of the row at once.  The reference's own value carries
    // the rounding of its nb-term sums (< 4e-4 relative for nb <= 6144 + the division), so outside alpha +- 1e-3 -- the margin
    // every bound of this file uses -- the decision is known; inside, `amb` sends the pair to the exact evaluation.
    // acc / amb: bits of M that certainly join / that need the exact evaluation (every lane gets the same words).
    auto fast_decide = [&](const uint4 info, const uint32_t M, const uint2* wr, uint32_t& acc, uint32_t& amb) {
        const uint32_t n = info.z - info.y;
        const float nc = sqrtf((float)info.w);
        const float lc = (float)info.x / nc;
        const bool mine = ((M >> lane) & 1u) && sq_s[lane] != 0 && info.w != 0;
        const uint32_t* rd = repd + (size_t)lane * p.nb;
        const float nr = nr_s[lane];
        float msum = 0.f;
        for (uint32_t i0 = 0; i0 < n; i0 += 8) {
            uint32_t v[8];
            float c[8];
#pragma unroll
            for (uint32_t u = 0; u < 8; ++u) {
                const uint2 pr = i0 + u < n ? wr[i0 + u] : make_uint2(0u, 0u);      // broadcast read
                const bool use = mine && (pr.y >> 31);
                v[u] = use ? __ldcg(rd + pr.x) : 0u;
                c[u] = (float)(pr.y & 0x7FFFFFFFu) / nc;
            }
#pragma unroll
            for (uint32_t u = 0; u < 8; ++u)
                if (v[u]) msum += fminf((float)v[u] / nr, c[u]);
        }
        const float sim = msum / (l_s[lane].x + lc - msum);
        const bool zero_case = ((M >> lane) & 1u) && !mine;               // a zero norm: the exact path knows the reference's rule
        const bool yes = mine && sim > p.alpha + 1e-3f;
        const bool no = mine && sim < bound;
        acc = __ballot_sync(0xffffffffu, yes);
        amb = __ballot_sync(0xffffffffu, zero_case || (mine && !yes && !no));
    };
    // The same decision for a row of more than 64 runs, read from global memory: lanes = runs (128 in flight), one candidate
    // representative after the other in ascending order, stopping at the first certain acceptance.  cand = the candidates that
    // are not certainly rejected, up to and including that acceptance; exact = false when one of them needs the exact evaluation.
    auto fast_decide_long = [&](const uint4 info, const uint32_t M, uint32_t& cand, bool& exact_needed) {
        const float nc = sqrtf((float)info.w);
        const float lc = (float)info.x / nc;
        cand = 0;
        exact_needed = false;
        for (uint32_t rest = M; rest; rest &= rest - 1) {
            const uint32_t r = __ffs(rest) - 1;
            if (sq_s[r] == 0 || info.w == 0) {                        // a zero norm: the exact path knows the reference's rule
                cand |= 1u << r;
                exact_needed = true;
                continue;
            }
            const uint32_t* rd = repd + (size_t)r * p.nb;
            const float nr = nr_s[r];
            float msum = 0.f;
            for (uint32_t j0 = info.y; j0 < info.z; j0 += 128) {
                uint2 pr4[4];
                uint32_t v4[4];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const uint32_t j = j0 + 32 * u + lane;
                    pr4[u] = j < info.z ? __ldg(p.enc_pair + j) : make_uint2(0u, 0u);
                }
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) v4[u] = (pr4[u].y >> 31) ? __ldcg(rd + pr4[u].x) : 0u;
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u)
                    if (v4[u]) msum += fminf((float)v4[u] / nr, (float)(pr4[u].y & 0x7FFFFFFFu) / nc);
            }
#pragma unroll
            for (int x = 1; x < 32; x <<= 1) msum += __shfl_xor_sync(0xffffffffu, msum, x);
            const float sim = msum / (l_s[r].x + lc - msum);
            if (sim < bound) continue;                                // certainly not
            cand |= 1u << r;
            if (sim > p.alpha + 1e-3f) break;                         // certainly: the row goes no further
            exact_needed = true;
        }
    };
    // Rows of 33..64 runs, one warp, runs staged in shared memory (`wr`, ascending blocks).  Same arithmetic as evaluate_pair;
    // a lane holds two runs (lane, lane + 32).  The run that comes first among those of a reference thread t leads it (found
    // with a match inside each half and a per-warp bitmap of the threads across the halves) and sums the thread's terms in
    // ascending block order, finding the row's later blocks of that thread by walking on through the staged runs.
    auto evaluate_medium = [&](const uint4 info, const uint32_t r, const uint2* wr, uint32_t* wbits) -> bool {
        const uint32_t s_cmp = info.w, s_rep = sq_s[r];
        if (s_rep == 0 && s_cmp == 0) return 1.0f > p.alpha;
        if (s_rep == 0 || s_cmp == 0) return 0.0f > p.alpha;
        const uint32_t* rd = repd + (size_t)r * p.nb;
        const float* pm = part_max + r * 1024;
        const float nr = nr_s[r];
        const uint32_t n = info.z - info.y;
        const float nc = sqrtf((float)s_cmp);
        float my_min = 0.f, my_max = lane < nw ? warp_max[r * 32 + lane] : 0.f;
        float pmn[2], pmx[2];
        uint32_t tt[2];
        bool lead[2];
#pragma unroll
        for (uint32_t h = 0; h < 2; ++h) {
            const uint32_t j = h * 32 + lane;
            const bool valid = j < n;
            const uint32_t blk = valid ? wr[j].x : 0u;
            const uint32_t t = mod_bd(blk);
            tt[h] = t;
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? t : 0xFFFF0000u + lane);
            const bool first = valid && lane == (uint32_t)(__ffs(peers) - 1);
            const bool ld = first && !((wbits[t >> 5] >> (t & 31u)) & 1u);       // no run of an earlier half belongs to t
            __syncwarp();
            if (ld) atomicOr(&wbits[t >> 5], 1u << (t & 31u));
            __syncwarp();
            lead[h] = ld;
            float pmin = 0.f, pmax = 0.f;
            if (ld) {
                uint32_t v[kStageTerms];
#pragma unroll
                for (uint32_t m = 0; m < kStageTerms; ++m) {
                    const uint32_t i = t + m * p.bd;
                    v[m] = i < p.nb ? __ldcg(rd + i) : 0u;
                }
                uint32_t nxt = j;
#pragma unroll
                for (uint32_t m = 0; m < kStageTerms; ++m) {
                    const uint32_t i = t + m * p.bd;
                    if (i < p.nb) {
                        const float a = (float)v[m] / nr;
                        while (nxt < n && wr[nxt].x < i) ++nxt;
                        if (nxt < n && wr[nxt].x == i) {
                            const float c = (float)(wr[nxt].y & 0x7FFFFFFFu) / nc;
                            pmin += fminf(a, c);
                            pmax += fmaxf(a, c);
                        } else {
                            pmax += a;
                        }
                    }
                }
            }
            pmn[h] = pmin;
            pmx[h] = pmax;
        }
        uint32_t touched = __reduce_or_sync(0xffffffffu, (lead[0] ? 1u << (tt[0] >> 5) : 0u) | (lead[1] ? 1u << (tt[1] >> 5) : 0u));
        while (touched) {
            const uint32_t w = __ffs(touched) - 1;
            touched &= touched - 1;
            float leaf_min = 0.f, leaf_max = pm[(w << 5) + lane];
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                uint32_t sel = __ballot_sync(0xffffffffu, lead[h] && (tt[h] >> 5) == w);
                while (sel) {
                    const int src = __ffs(sel) - 1;
                    sel &= sel - 1;
                    const uint32_t tl = __shfl_sync(0xffffffffu, tt[h] & 31u, src);
                    const float vmin = __shfl_sync(0xffffffffu, pmn[h], src);
                    const float vmax = __shfl_sync(0xffffffffu, pmx[h], src);
                    if (lane == tl) {
                        leaf_min = vmin;
                        leaf_max = vmax;
                    }
                }
            }
#pragma unroll
            for (int x = 1; x < 32; x <<= 1) {
                leaf_min += __shfl_xor_sync(0xffffffffu, leaf_min, x);
                leaf_max += __shfl_xor_sync(0xffffffffu, leaf_max, x);
            }
            if (lane == w) {
                my_min = leaf_min;
                my_max = leaf_max;
            }
        }
        __syncwarp();
        wbits[lane] = 0;
        __syncwarp();
        for (uint32_t stride = p.first_stride; stride >= 1; stride >>= 1) {
            const float tmin = __shfl_down_sync(0xffffffffu, my_min, stride);
            const float tmax = __shfl_down_sync(0xffffffffu, my_max, stride);
            if (lane < stride && lane + stride < 32) {
                my_min += tmin;
                my_max += tmax;
            }
        }
        const float sim = __shfl_sync(0xffffffffu, my_min, 0) / __shfl_sync(0xffffffffu, my_max, 0);
        return sim > p.alpha;
    };
    // Rows of more than 32 runs: the row is expanded ONCE into the CTA's dense scratch (16-bit counts: a count never exceeds
    // the block size) and up to 32 warps evaluate it against 32 representatives at the same time -- every term of every
    // touched reference thread is one shared-memory read and one coalesced read of the dense representative, no search.
    auto expand_row = [&](const uint4 info) {
        uint32_t touched = 0;
        for (uint32_t j = info.y + tid; j < info.z; j += kClusterThreads) {
            const uint32_t blk = __ldg(p.enc_blk + j);
            scratch[blk] = (uint16_t)__ldg(p.counts + j);
            touched |= 1u << (mod_bd(blk) >> 5);
        }
        touched = __reduce_or_sync(0xffffffffu, touched);
        if (lane == 0 && touched) atomicOr(&s_touched, touched);
    };
    auto clear_row = [&](const uint4) {
        for (uint32_t i = tid; i < (nbp >> 1); i += kClusterThreads) reinterpret_cast<uint32_t*>(scratch)[i] = 0;     // 3 words per thread: cheaper than reading the row again
        if (tid == 0) s_touched = 0;
    };
    // The whole CTA, after expand_row + barrier: does the row in the scratch join one of the representatives `mk`?  CTA warp w
    // plays REFERENCE warp w for every candidate representative: the row's terms of reference thread t = 32 w + lane are read
    // from the scratch once; per representative the thread's sums are recomputed from the dense counts only where the row
    // has a block (others: the representative's own partial sum), two representatives per iteration so that their loads
    // overlap; butterfly; the per-warp pairs go to s_res.  After a barrier one warp per representative runs the reference's
    // tree over the 32 pairs and records an acceptance as key | r in s_first.  Ends with a barrier.
    auto scratch_pass = [&](const uint4 info, const uint32_t mk, const uint32_t key, uint32_t* first) {
        const uint32_t touched = s_touched;
        const float nc = sqrtf((float)info.w);
        if (wid < nw && ((touched >> wid) & 1u)) {
            const uint32_t t = (wid << 5) + lane;
            float cf[kStageTerms];                              // the row's normalised count per term, 0 = no block
            bool mine = false;
#pragma unroll
            for (uint32_t m = 0; m < kStageTerms; ++m) {
                const uint32_t i = t + m * p.bd;
                const uint32_t cnt = i < p.nb ? scratch[i] : 0u;
                cf[m] = cnt ? (float)cnt / nc : 0.f;
                mine |= cnt != 0;
            }
            auto one = [&](const uint32_t r, float& amin, float& amax) {
                const uint32_t* rd = repd + (size_t)r * p.nb;
                const float nr = nr_s[r];
                amin = 0.f;
                if (!mine) { amax = part_max[r * 1024 + t]; return; }
                uint32_t v[kStageTerms];
#pragma unroll
                for (uint32_t m = 0; m < kStageTerms; ++m) {
                    const uint32_t i = t + m * p.bd;
                    v[m] = i < p.nb ? __ldcg(rd + i) : 0u;
                }
                amax = 0.f;
#pragma unroll
                for (uint32_t m = 0; m < kStageTerms; ++m) {
                    if (t + m * p.bd < p.nb) {
                        const float a = (float)v[m] / nr;
                        if (cf[m] != 0.f) {
                            amin += fminf(a, cf[m]);
                            amax += fmaxf(a, cf[m]);
                        } else {
                            amax += a;
                        }
                    }
                }
            };
            uint32_t rest = mk;
            while (rest) {
                const uint32_t ra = __ffs(rest) - 1;
                rest &= rest - 1;
                const uint32_t rb = rest ? __ffs(rest) - 1 : ra;
                rest &= rest - 1;                               // (0 & anything = 0)
                float amin, amax, bmin, bmax;
                one(ra, amin, amax);
                one(rb, bmin, bmax);
#pragma unroll
                for (int x = 1; x < 32; x <<= 1) {
                    amin += __shfl_xor_sync(0xffffffffu, amin, x);
                    amax += __shfl_xor_sync(0xffffffffu, amax, x);
                    bmin += __shfl_xor_sync(0xffffffffu, bmin, x);
                    bmax += __shfl_xor_sync(0xffffffffu, bmax, x);
                }
                if (lane == 0) {
                    s_res[ra * 32 + wid] = make_float2(amin, amax);
                    s_res[rb * 32 + wid] = make_float2(bmin, bmax);
                }
            }
        }
        __syncthreads();
        if (wid < __popc(mk)) {
            const uint32_t r = __fns(mk, 0, (int)wid + 1);
            const uint32_t s_rep = sq_s[r], s_cmp = info.w;
            bool joins;
            if (s_rep == 0 && s_cmp == 0) joins = 1.0f > p.alpha;
            else if (s_rep == 0 || s_cmp == 0) joins = 0.0f > p.alpha;
            else {
                float my_min = 0.f, my_max = 0.f;               // lane = reference warp
                if (lane < nw) {
                    if ((touched >> lane) & 1u) {
                        const float2 v = s_res[r * 32 + lane];
                        my_min = v.x;
                        my_max = v.y;
                    } else {
                        my_max = warp_max[r * 32 + lane];
                    }
                }
                for (uint32_t stride = p.first_stride; stride >= 1; stride >>= 1) {
                    const float tmin = __shfl_down_sync(0xffffffffu, my_min, stride);
                    const float tmax = __shfl_down_sync(0xffffffffu, my_max, stride);
                    if (lane < stride && lane + stride < 32) {
                        my_min += tmin;
                        my_max += tmax;
                    }
                }
                const float sim = __shfl_sync(0xffffffffu, my_min, 0) / __shfl_sync(0xffffffffu, my_max, 0);
                joins = sim > p.alpha;
            }
            if (joins && lane == 0) atomicMin(first, key | r);
        }
        __syncthreads();
    };
    // The whole CTA, row expanded in the scratch (barrier done): the shared-nnz bound of warp_shared_mask without touching
    // global memory -- every thread looks at its column blocks.  Two barriers; every thread returns the same mask.
    auto shared_mask_scratch = [&](const uint4 info, uint32_t alive) -> uint32_t {
        if (tid < kStageReps) s_sh[tid] = 0;
        __syncthreads();
        for (uint32_t blk = tid; blk < p.nb; blk += kClusterThreads) {
            const uint32_t cnt = scratch[blk];
            if (cnt && ((p.kept_mask >> (mod_bd(blk) >> 5)) & 1u)) {
                uint32_t m = slots[blk] & alive;
                while (m) {
                    atomicAdd(&s_sh[__ffs(m) - 1], cnt);
                    m &= m - 1;
                }
            }
        }
        __syncthreads();
        const uint32_t sh = s_sh[lane];
        const float nc = sqrtf((float)info.w);
        const bool ok = sh != 0 && !((float)sh < bound * (float)info.x) && keep_by_union(sh, nc, (float)info.x / nc, lane);
        return __ballot_sync(0xffffffffu, ok) & alive;
    };
    // A row joins representative r: absorb + refresh in one, with the integer sums updated from the row's blocks alone
    // (sum of squares += (old + cnt)^2 - old^2 over the kept blocks, wrapping like the reference's) and ONE dense pass.
    auto join_absorb_refresh = [&](const uint4 info, uint32_t r, uint32_t nrep_now, const uint2* sruns) {
        uint32_t* rd = repd + (size_t)r * p.nb;
        uint32_t dsq = 0, dtot = 0;
        for (uint32_t j = info.y + tid; j < info.z; j += kClusterThreads) {
            const uint2 pr = sruns ? sruns[tid] : __ldg(p.enc_pair + j);
            const uint32_t cnt = pr.y & 0x7FFFFFFFu;
            const uint32_t old = __ldcg(rd + pr.x);
            rd[pr.x] = old + cnt;
            atomicOr(&slots[pr.x], 1u << r);
            if (pr.y >> 31) {
                dsq += (old + cnt) * (old + cnt) - old * old;
                dtot += cnt;
            }
        }
        dsq = __reduce_add_sync(0xffffffffu, dsq);
        dtot = __reduce_add_sync(0xffffffffu, dtot);
        if (lane == 0 && (dsq | dtot)) { atomicAdd(&s_red[0], dsq); atomicAdd(&s_red[1], dtot); }     // s_red is zero between uses
        if (tid == 0 && single_s[r] != kNone) {                      // still a single block?
            const uint2 pr0 = sruns ? sruns[0] : __ldg(p.enc_pair + info.y);
            if (info.z - info.y == 1 && pr0.x == single_s[r]) sc_s[r] += pr0.y & 0x7FFFFFFFu;
            else single_s[r] = kNone;
        }
        __syncthreads();
        const uint32_t sqr = sq_s[r] + s_red[0], ttr = tot_s[r] + s_red[1];
        const float nr = sqrtf((float)sqr);
        float mx = 0.f;
        if (tid < p.bd) {
            for (uint32_t i = tid; i < p.nb; i += p.bd) mx += (float)__ldcg(rd + i) / nr;
            part_max[r * 1024 + tid] = mx;
        }
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) mx += __shfl_xor_sync(0xffffffffu, mx, w);
        if (lane == 0) warp_max[r * 32 + wid] = tid < p.bd ? mx : 0.f;
        __syncthreads();
        if (tid == 0) {
            sq_s[r] = sqr;
            tot_s[r] = ttr;
            nr_s[r] = nr;
            const float l1 = sqr ? (float)ttr / nr : 0.f;
            l_s[r] = make_float2(l1, bound * l1);
            if (sqr) atomicOr(&s_nz, 1u << r);
            s_red[0] = 0;
            s_red[1] = 0;
        }
        __syncthreads();
        if (wid == 0) {
            const bool on = lane < nrep_now && ((s_nz >> lane) & 1u);
            float lo = on ? l_s[lane].y : INFINITY, hi = on ? l_s[lane].x : 0.f;
#pragma unroll
            for (int w = 1; w < 32; w <<= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, w));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, w));
            }
            if (lane == 0) { s_blr_min = lo; s_lr_max = hi; }
        }
        __syncthreads();
    };
    // A row founds representative r: everything refresh() derives, computed by WARP 0 from the row's runs alone (the dense
    // sums over a single row are sums over its runs: zero terms add 0.f).  part_max / warp_max of r are zero on entry (stage
    // start).  Also refreshes the size-bound summary for `nrep_now` representatives.  The caller's next barrier publishes it.
    auto found_sparse = [&](const uint4 info, uint32_t r, uint32_t nrep_now, const uint2* sruns) {
        float* pm = part_max + r * 1024;
        const float nr = sqrtf((float)info.w);
        uint32_t touched = 0;
        for (uint32_t j0 = info.y; j0 < info.z; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool valid = j < info.z;
            const uint2 spr = (valid && sruns) ? sruns[j - info.y] : make_uint2(0u, 0u);
            const uint32_t blk = valid ? (sruns ? spr.x : __ldg(p.enc_blk + j)) : 0u;
            const uint32_t cnt = valid ? (sruns ? (spr.y & 0x7FFFFFFFu) : __ldg(p.counts + j)) : 0u;
            const uint32_t t = mod_bd(blk);
            float* fv = s_fv + wid * 32;
            fv[lane] = (float)cnt / nr;
            __syncwarp();
            // runs of the same reference thread (blk = t, t + bd, ...) are added in ascending block order = ascending lane
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? t : 0xFFFF0000u + lane);
            if (valid && lane == (uint32_t)(__ffs(peers) - 1)) {
                float acc = pm[t];
                for (uint32_t rest = peers; rest; rest &= rest - 1) acc += fv[__ffs(rest) - 1];
                pm[t] = acc;
                touched |= 1u << (t >> 5);
            }
            __syncwarp();
        }
        touched = __reduce_or_sync(0xffffffffu, touched);
        while (touched) {
            const uint32_t w = __ffs(touched) - 1;
            touched &= touched - 1;
            float mx = pm[(w << 5) + lane];
#pragma unroll
            for (int x = 1; x < 32; x <<= 1) mx += __shfl_xor_sync(0xffffffffu, mx, x);
            if (lane == 0) warp_max[r * 32 + w] = mx;
        }
        if (lane == 0) {
            sq_s[r] = info.w;
            tot_s[r] = info.x;
            nr_s[r] = nr;
            const float l1 = info.w ? (float)info.x / nr : 0.f;
            l_s[r] = make_float2(l1, bound * l1);
            if (info.w) atomicOr(&s_nz, 1u << r);
        }
        if (nrep_now == 0) return;                                   // (several founders at once: the caller refreshes the summary)
        __syncwarp();
        const bool on = lane < nrep_now && ((s_nz >> lane) & 1u);
        float lo = on ? l_s[lane].y : INFINITY, hi = on ? l_s[lane].x : 0.f;
#pragma unroll
        for (int x = 1; x < 32; x <<= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, x));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, x));
        }
        if (lane == 0) { s_blr_min = lo; s_lr_max = hi; }
    };
    unsigned long long s_ctrl_copy = 0;
    auto poll = [&](uint32_t id, uint32_t have) -> bool {
        if (tid == 0) {
            const unsigned long long* cw = p.ctrl + (id % p.num_slots);
            uint32_t spins = 0;
            unsigned long long v, t0 = 0;
            uint32_t stop = 0;
            for (;;) {
                v = ld_acquire_u64(cw);
                if ((uint32_t)(v >> 33) == id && ((uint32_t)((v >> 1) & 0xFFFFFFFFu) > have || (v & 1ull))) break;
                if (((volatile uint32_t*)p.status)[1] | ((volatile uint32_t*)p.status)[2]) { stop = 1; break; }
                if ((++spins & 1023u) == 0) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    if (now - t0 > 120ull * 1000000000ull) { atomicExch(p.status + 2, 1u); stop = 1; break; }
                }
                __nanosleep(spins < 64 ? 20 : 200);
            }
            s_ctrl = v;
            s_stop = stop;
        }
        __syncthreads();
        const bool ok = s_stop == 0;
        const unsigned long long v = s_ctrl;
        __syncthreads();
        s_ctrl_copy = v;
        return ok;
    };

    if (tid < 2) s_first[tid] = kNone;
    if (tid == 0) { s_touched = 0; s_red[0] = 0; s_red[1] = 0; s_anyq = 0; }
    s_wbits[tid] = 0;
    for (uint32_t i = tid; i < (nbp >> 1); i += kClusterThreads) reinterpret_cast<uint32_t*>(scratch)[i] = 0;
    uint32_t par_a = 0, par_b = 0, par_c = 0;
    auto now_ns = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    unsigned long long tr_steps = 0, tr_cand = 0, tr_joins = 0, tr_polls = 0, tr_events = 0, tr_found = 0;
    unsigned long long tr_long = 0, tr_poll_cyc = 0, tr_found_cyc = 0, tr_event_cyc = 0, tr_long_cyc = 0;
    const long long tr_begin = clock64();
    auto flush_trace = [&]() {
        if (tid == 0 && p.trace) {
            atomicAdd(p.trace + 0, tr_steps); atomicAdd(p.trace + 1, tr_cand); atomicAdd(p.trace + 2, tr_joins);
            atomicAdd(p.trace + 3, tr_polls); atomicAdd(p.trace + 4, tr_events); atomicAdd(p.trace + 5, tr_found);
            atomicAdd(p.trace + 6, tr_long); atomicAdd(p.trace + 7, (unsigned long long)(clock64() - tr_begin));
            atomicAdd(p.trace + 8, tr_poll_cyc); atomicAdd(p.trace + 9, tr_found_cyc); atomicAdd(p.trace + 10, tr_event_cyc);
            atomicAdd(p.trace + 11, tr_long_cyc);
        }
    };
    for (uint32_t s = blockIdx.x + 1;; s += gridDim.x) {
        const uint32_t* in = p.lists + (size_t)(s % p.num_slots) * p.list_cap;
        uint32_t* out = p.lists + (size_t)((s + 1) % p.num_slots) * p.list_cap;
        unsigned long long* out_ctrl = p.ctrl + ((s + 1) % p.num_slots);
        // reset the stage's state while the input is on its way
        for (uint32_t i = tid; i < nbp; i += kClusterThreads) slots[i] = 0;
        for (uint32_t i = tid; i < kStageReps * p.nb; i += kClusterThreads) repd[i] = 0;
        for (uint32_t i = tid; i < kStageReps * (1024 + 32); i += kClusterThreads) part_max[i] = 0.f;     // part_max and warp_max
        if (tid == 0) { s_nz = 0; s_blr_min = INFINITY; s_lr_max = 0.f; }
        if (tid < kStageReps) { single_s[tid] = kNone; sc_s[tid] = 0; s_add[tid] = 0; }
        ++tr_polls;
        {
            const long long t0 = clock64();
            const bool ok = poll(s, 0);
            tr_poll_cyc += clock64() - t0;
            if (!ok) { flush_trace(); return; }
        }
        uint32_t avail = (uint32_t)((s_ctrl_copy >> 1) & 0xFFFFFFFFu);
        bool in_done = (s_ctrl_copy & 1ull) != 0;
        if (avail == 0) {                       // the previous stage forwarded nothing (it has recorded the cluster count): the run is over
            if (tid == 0) atomicExch(p.status + 1, 1u);
            flush_trace();
            return;
        }
        if (tid == 0) st_release_u64(out_ctrl, make_ctrl(s + 1, 0, 0));
        if (tid == 0 && p.trace_ts && s <= p.M) { p.trace_ts[3 * s] = now_ns(); p.trace_ts[3 * s + 1] = 0; }
        const uint32_t base = (s - 1) * kStageReps + 1;      // cluster id of representative 0
        uint32_t cursor = 0, produced = 0, published = 0, nrep = 0;
        uint32_t fb_base = 0, fb_n = 0;                      // founding: the rows fetched into s_fpos / s_finfo / s_fruns
        uint32_t fresh_steps = 0;                            // streaming steps since the 32nd representative
        bool careful = false;                                // 32 representatives, but joins are coming in bursts: rounds of 32 rows
        uint32_t dbg_rounds = 0, dbg_rj_careful = 0, dbg_rj_stream = 0, dbg_cj = 0, dbg_steps = 0, dbg_solo = 0;   // (debug trace)
        long long dbg_cyc_careful = 0, dbg_cyc_stream = 0;
        uint32_t calm_rows = 0;                              // rows settled since the last join
        for (;;) {
            if (cursor >= avail) {
                if (in_done) break;
                if (produced != published) {                 // nothing to do until more rows arrive: do not sit on the ones already decided
                    __syncthreads();                           // (at the end of the order the rows trickle through 148 stages one at a time)
                    if (tid == 0) st_release_u64(out_ctrl, make_ctrl(s + 1, produced, 0));
                    published = produced;
                }
                ++tr_polls;
                {
                    const long long t0 = clock64();
                    const bool ok = poll(s, cursor);
                    tr_poll_cyc += clock64() - t0;
                    if (!ok) { flush_trace(); return; }
                }
                avail = (uint32_t)((s_ctrl_copy >> 1) & 0xFFFFFFFFu);
                in_done = (s_ctrl_copy & 1ull) != 0;
                continue;
            }
            if (nrep < kStageReps || careful) {
                // ---- careful mode: rounds over a batch of 32 rows held in shared memory ----
                // Used while the stage has fewer than 32 representatives (a row that joins nothing founds the next one, and nothing is
                // forwarded before the 32nd: this is the critical path of the whole pipeline) and after a join in streaming mode (on
                // graphs the joins come in bursts -- every row that holds a hub block joins that block's cluster -- and a join
                // invalidates whatever was computed for the rows behind it: 1024-row steps would be thrown away 30 times over).
                // The batch's positions, pos_info and block lists (rows of at most 64 runs) are fetched together.  A round: warp w
                // decides row w of what is left of the batch against the current representatives; the first row that changes the
                // stage (a join that is not a certain one, a founder, a row that needs the CTA-wide scratch) ends the round, everything
                // before it is settled (certain joins, forwarded rows), the event is applied, and the next round re-decides only the
                // rows behind it, out of shared memory.
                const long long tf0 = clock64();
                if (cursor >= fb_base + fb_n) {
                    fb_base = cursor;
                    fb_n = min(32u, avail - cursor);
                    if (tid < fb_n) {
                        const uint32_t ps = __ldcg(in + cursor + tid);
                        s_fpos[tid] = ps;
                        s_finfo[tid] = __ldg(p.pos_info + ps);
                    }
                    __syncthreads();
                    if (wid < fb_n) {
                        const uint4 fi = s_finfo[wid];
                        const uint32_t fn = fi.z - fi.y;
                        if (fn <= 64) {
                            if (lane < fn) s_runs64[wid * 64 + lane] = __ldg(p.enc_pair + fi.y + lane);
                            if (lane + 32 < fn) s_runs64[wid * 64 + lane + 32] = __ldg(p.enc_pair + fi.y + lane + 32);
                        }
                    }
                    __syncthreads();
                }
                const uint32_t fk = cursor - fb_base;               // first undecided row of the batch
                const uint32_t nrows = fb_n - fk;
                constexpr uint32_t kReject = 0, kCertain = 1, kAccept = 2, kSolo = 3;
                if (wid < nrows) {
                    const uint32_t k = fk + wid;
                    const uint4 info = s_finfo[k];
                    const uint32_t fn = info.z - info.y;
                    const uint2* sruns = fn <= 64 ? s_runs64 + k * 64 : nullptr;
                    uint32_t M = nrep ? size_mask(info, nrep) : 0u;
                    if (M && info.w && fn <= 64) M = warp_shared_mask(info, M, sruns);     // (longer rows: the CTA-wide scratch does it, once)
                    uint32_t code = kReject;
                    if (M) {
                        if (fn > 64) {
                            code = kSolo;
                        } else {
                            bool certain = false;
                            if (cj_on && !(M & (M - 1)) && fn == 1) {
                                const uint32_t r = __ffs(M) - 1;
                                const uint2 pr = sruns[0];
                                certain = single_s[r] != kNone && pr.x == single_s[r] && (pr.y >> 31) && (pr.y & 0x7FFFFFFFu) <= 32u && sc_s[r] <= 32767u;
                                if (certain) code = kCertain | (r << 8);
                            }
                            if (!certain) {
                                uint32_t acc, amb;
                                fast_decide(info, M, sruns, acc, amb);
                                uint32_t mk = acc | amb;
                                while (mk) {
                                    const uint32_t r = __ffs(mk) - 1;
                                    mk &= mk - 1;
                                    if (((acc >> r) & 1u) || (fn <= 32 ? evaluate_pair(info, r, sruns) : evaluate_medium(info, r, sruns, s_wbits + wid * 32))) {
                                        code = kAccept | (r << 8);
                                        break;
                                    }
                                }
                            }
                        }
                    }
                    if (lane == 0) s_out[par_c][k] = code;
                }
                if (tid < 32) s_pair[tid] = 0;
                __syncthreads();                                   // R1: every row of the round decided
                if (tid == 0 && produced != published) st_release_u64(out_ctrl, make_ctrl(s + 1, produced, 0));
                published = produced;
                const uint32_t my_code = lane < nrows ? s_out[par_c][fk + lane] : kReject;
                par_c ^= 1;
                const uint32_t my_type = my_code & 0xFFu;
                const bool founding = nrep < kStageReps;
                const uint32_t my_n = lane < nrows ? s_finfo[fk + lane].z - s_finfo[fk + lane].y : 0u;
                const uint32_t stopm = __ballot_sync(0xffffffffu, lane < nrows && (my_type == kAccept || my_type == kSolo));
                const uint32_t e_stop = stopm ? __ffs(stopm) - 1 : nrows;
                const uint32_t rejm_all = __ballot_sync(0xffffffffu, lane < e_stop && my_type == kReject);
                uint32_t founders = 0, e = e_stop;
                if (founding && rejm_all) {
                    // ---- founders, several per round ----
                    // Every row that joins nothing founds a representative -- but the second must also be tested against the first.
                    // Thread (i, j) of the CTA tests row j of the batch against a cluster made of row i alone with the sparse
                    // similarity (fast_decide's formula over two block lists in shared memory); rows that pass nothing, in order,
                    // are founded together, a warp each.  A row that might join an earlier founder of the round ends it and is
                    // decided in the next round.
                    {
                        const uint32_t i = wid, j = lane;            // thread (i, j), i < j
                        bool maybe = false;
                        if (i < j && j < e_stop && ((rejm_all >> i) & 1u) && ((rejm_all >> j) & 1u)) {
                            const uint4 ii = s_finfo[fk + i], ij = s_finfo[fk + j];
                            const uint32_t ni = ii.z - ii.y, nj = ij.z - ij.y;
                            if (ni > 64 || nj > 64 || ii.w == 0 || ij.w == 0) {
                                maybe = true;
                            } else {
                                const uint2* ri = s_runs64 + (fk + i) * 64;
                                const uint2* rj = s_runs64 + (fk + j) * 64;
                                const float nci = sqrtf((float)ii.w), ncj = sqrtf((float)ij.w);
                                float msum = 0.f;
                                uint32_t a = 0, b = 0;
                                while (a < ni && b < nj) {
                                    const uint2 pa = ri[a], pb = rj[b];
                                    if (pa.x == pb.x) {
                                        if (pa.y >> 31) msum += fminf((float)(pa.y & 0x7FFFFFFFu) / nci, (float)(pb.y & 0x7FFFFFFFu) / ncj);
                                        ++a;
                                        ++b;
                                    } else if (pa.x < pb.x) {
                                        ++a;
                                    } else {
                                        ++b;
                                    }
                                }
                                const float sim = msum / ((float)ii.x / nci + (float)ij.x / ncj - msum);
                                maybe = !(sim < bound);
                            }
                        }
                        const uint32_t mm = __ballot_sync(0xffffffffu, maybe);     // lanes = j, this warp = i
                        for (uint32_t rest = mm; rest; rest &= rest - 1)
                            if (lane == (uint32_t)(__ffs(rest) - 1)) atomicOr(&s_pair[lane], 1u << i);
                    }
                    __syncthreads();
                    uint32_t count = 0;
                    e = e_stop;
                    for (uint32_t j = 0; j < e_stop; ++j) {          // (every thread runs the same scan)
                        if (!((rejm_all >> j) & 1u)) continue;       // a certain join: stays valid, its representative comes first
                        const uint32_t nj = __shfl_sync(0xffffffffu, my_n, j);
                        if ((s_pair[j] & founders) || nrep + count == kStageReps || (nj > 64 && founders)) { e = j; break; }
                        founders |= 1u << j;
                        ++count;
                        if (nj > 64) { e = j + 1; break; }           // (founded through the dense refresh below, alone)
                    }
                }
                // settle the rows before it
                const uint32_t cjm = __ballot_sync(0xffffffffu, lane < e && my_type == kCertain);
                const uint32_t rjm = __ballot_sync(0xffffffffu, lane < e && my_type == kReject && !founding);     // (only with 32 representatives)
                if (wid == 0 && lane < e) {
                    const uint32_t ps = s_fpos[fk + lane];
                    if (my_type == kCertain) {
                        const uint32_t r = my_code >> 8;
                        atomicAdd(&s_add[r], s_runs64[(fk + lane) * 64].y & 0x7FFFFFFFu);
                        p.cluster_ids[ps] = base + r;
                    } else if (!founding) {
                        out[produced + __popc(rjm & ((1u << lane) - 1u))] = ps;
                    }
                }
                if (founders) {
                    // ---- found them: warp w takes the w-th, representative nrep + w ----
                    const uint32_t count = __popc(founders);
                    bool big_last = false;
                    if (wid < count) {
                        const uint32_t j = __fns(founders, 0, (int)wid + 1);
                        const uint32_t r = nrep + wid;
                        const uint4 info = s_finfo[fk + j];
                        const uint32_t fn = info.z - info.y;
                        if (lane == 0) p.cluster_ids[s_fpos[fk + j]] = base + r;
                        if (fn <= 64) {
                            const uint2* sruns = s_runs64 + (fk + j) * 64;
                            for (uint32_t x = lane; x < fn; x += 32) {     // the representative was all zero: plain stores
                                const uint2 pr = sruns[x];
                                repd[(size_t)r * p.nb + pr.x] = pr.y & 0x7FFFFFFFu;
                                atomicOr(&slots[pr.x], 1u << r);
                            }
                            found_sparse(info, r, 0u, sruns);
                            if (lane == 0 && fn == 1 && (sruns[0].y >> 31)) {   // a single kept block: certain joins possible
                                single_s[r] = sruns[0].x;
                                sc_s[r] = sruns[0].y & 0x7FFFFFFFu;
                            }
                        }
                    }
                    {   // (a founder of more than 64 runs is the last of its round)
                        const uint32_t jl = 31 - __clz(founders);
                        big_last = __shfl_sync(0xffffffffu, my_n, jl) > 64;
                        if (big_last) {
                            const uint32_t r = nrep + count - 1;
                            absorb(s_finfo[fk + jl], r);
                            refresh(r, nrep + count);
                        }
                    }
                    nrep += count;
                    tr_found += count;
                    __syncthreads();
                    if (wid == 0) {                                // the size-bound summary over all representatives
                        const bool on = lane < nrep && ((s_nz >> lane) & 1u);
                        float lo = on ? l_s[lane].y : INFINITY, hi = on ? l_s[lane].x : 0.f;
#pragma unroll
                        for (int x = 1; x < 32; x <<= 1) {
                            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, x));
                            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, x));
                        }
                        if (lane == 0) { s_blr_min = lo; s_lr_max = hi; }
                    }
                    __syncthreads();
                    if (nrep == kStageReps) {
                        fresh_steps = 0;
                        careful = false;
                        if (tid == 0 && p.trace_ts && s <= p.M) p.trace_ts[3 * s + 1] = now_ns();
                    }
                }
                produced += __popc(rjm);
                calm_rows += e;
                tr_joins += __popc(cjm);
                tr_cand += e;
                if (cjm) {                                         // (uniform)
                    __syncthreads();
                    if (tid < kStageReps && s_add[tid]) {
                        const uint32_t c2 = sc_s[tid] + s_add[tid];
                        s_add[tid] = 0;
                        sc_s[tid] = c2;
                        repd[(size_t)tid * p.nb + single_s[tid]] = c2;
                        sq_s[tid] = c2 * c2;
                        tot_s[tid] = c2;
                        nr_s[tid] = sqrtf((float)(c2 * c2));
                    }
                    __syncthreads();
                }
                ++tr_steps;
                ++dbg_rounds;
                dbg_cj += __popc(cjm);
                if (e == nrows) dbg_cyc_careful += clock64() - tf0;
                if (e == nrows) {                                  // nothing changed the stage: the batch is used up
                    cursor = fb_base + fb_n;
                    if (!founding && calm_rows >= 64) careful = false;
                    tr_found_cyc += clock64() - tf0;
                    continue;
                }
                if (founders || (founding && e < e_stop)) {        // founders were made (or the stage is full): the row at e is decided next round
                    cursor = fb_base + fk + e;
                    tr_found_cyc += clock64() - tf0;
                    dbg_cyc_careful += clock64() - tf0;
                    continue;
                }
                // ---- the event row ----
                const uint32_t k = fk + e;
                const uint32_t ecode = __shfl_sync(0xffffffffu, my_code, e);
                const uint32_t pos = s_fpos[k];
                const uint4 info = s_finfo[k];
                const uint32_t fn = info.z - info.y;
                const bool big = fn > 64;
                const uint2* sruns = big ? nullptr : s_runs64 + k * 64;
                uint32_t jr = kNone;
                bool expanded = false;
                if ((ecode & 0xFFu) == kAccept) {
                    jr = ecode >> 8;
                } else if ((ecode & 0xFFu) == kSolo) {
                    // a row of more than 64 runs with candidates left: the CTA-wide scratch
                    uint32_t M = size_mask(info, nrep);
                    expand_row(info);
                    expanded = true;
                    __syncthreads();
                    if (M && info.w) M = shared_mask_scratch(info, M);
                    if (M) {
                        scratch_pass(info, M, 0u, &s_first[par_b]);
                        jr = s_first[par_b];
                        par_b ^= 1;
                        if (tid == 0) s_first[par_b] = kNone;      // visible after the barrier(s) below
                    }
                    clear_row(info);                               // closed by the barrier(s) below
                }
                if ((ecode & 0xFFu) == kSolo) ++dbg_solo;
                if (jr != kNone) {
                    if (tid == 0) p.cluster_ids[pos] = base + jr;
                    ++tr_joins;
                    ++dbg_rj_careful;
                    join_absorb_refresh(info, jr, nrep, sruns);
                    calm_rows = 0;
                } else if (founding) {
                    const uint32_t r = nrep;
                    if (tid == 0) p.cluster_ids[pos] = base + r;
                    ++nrep;
                    ++tr_found;
                    if (!big) {
                        if (tid >= 32 && tid - 32 < fn) {          // the representative was all zero: plain stores
                            const uint2 pr = sruns[tid - 32];
                            repd[(size_t)r * p.nb + pr.x] = pr.y & 0x7FFFFFFFu;
                            atomicOr(&slots[pr.x], 1u << r);
                        }
                        if (wid == 0) found_sparse(info, r, nrep, sruns);
                        if (tid == 32 && fn == 1 && (sruns[0].y >> 31)) {      // a single kept block: certain joins possible
                            single_s[r] = sruns[0].x;
                            sc_s[r] = sruns[0].y & 0x7FFFFFFFu;
                        }
                        __syncthreads();
                    } else {
                        absorb(info, r);
                        refresh(r, nrep);
                    }
                    if (nrep == kStageReps) {
                        fresh_steps = 0;
                        careful = false;
                        if (tid == 0 && p.trace_ts && s <= p.M) p.trace_ts[3 * s + 1] = now_ns();
                    }
                } else {
                    // (a kSolo row that joined nothing, 32 representatives): it moves on
                    if (tid == 0) out[produced] = pos;
                    produced += 1;
                    calm_rows += 1;
                    __syncthreads();                               // closes clear_row
                }
                (void)expanded;
                cursor = fb_base + k + 1;
                ++tr_cand;
                tr_found_cyc += clock64() - tf0;
                dbg_cyc_careful += clock64() - tf0;
                continue;
            }
            // ---- streaming step: up to 1024 rows, one per thread, against all 32 representatives ----
            // (measured: 8 per warp -> 3.4 s instead of 3.0 s at 2^20 rows, 23.6 against 23.8 s at 2^22)
            constexpr uint32_t kCpwB = BSMR_STAGE_CPW;                       // event rows a warp evaluates per step (measured: 11 % of the rows of a step are events)
            // the first steps after the founding are short and publish at once: the next stage is waiting for its first rows
            const bool fresh = fresh_steps < 3;
            ++fresh_steps;
            fb_n = 0;                                           // (the warps stage rows in s_runs64 below: the careful-mode batch is gone)
            const long long ts0 = clock64();
            ++dbg_steps;
            const uint32_t take = min(avail - cursor, fresh ? 128u : (uint32_t)kClusterThreads);
            ++tr_steps;
            tr_cand += take;
            uint32_t pos1 = kNone, M = 0, cj_r = 0, cj_cnt = 0;
            bool is_long = false, cj = false;
            if (tid < take) {
                pos1 = __ldcg(in + cursor + tid);
                const uint4 info = __ldg(p.pos_info + pos1);
                M = bounds_mask(info, kStageReps, is_long);
                if (M && !is_long && info.w != 0 && __popc(M) <= 8 && !(cj_on && info.z - info.y == 1)) {
                    // Fast decision by the THREAD for the candidates the bounds left, two per walk over the row (see fast_decide: the
                    // sparse min-sum gives the similarity to ~1e-6): measured on R-MAT graphs 11-21 % of the rows of a step survive the
                    // bounds and 99.98 % of those are then rejected by a warp at four dependent round trips each -- 46 % of the kernel
                    // at 2^22 rows.  Here 1024 threads do it at once; what is left for the warps is what may actually join.
                    const float nc = sqrtf((float)info.w);
                    const float lc = (float)info.x / nc;
                    for (uint32_t rest = M; rest;) {
                        const uint32_t r0 = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const bool two = rest != 0;
                        const uint32_t r1 = two ? __ffs(rest) - 1 : r0;
                        rest &= rest - 1;                       // (0 & anything = 0)
                        const uint32_t* rd0 = repd + (size_t)r0 * p.nb;
                        const uint32_t* rd1 = repd + (size_t)r1 * p.nb;
                        const float nr0 = nr_s[r0], nr1 = nr_s[r1];
                        float m0 = 0.f, m1 = 0.f;
                        for (uint32_t j = info.y; j < info.z; j += 4) {
                            uint2 pr[4];
                            uint32_t v0[4], v1[4];
#pragma unroll
                            for (uint32_t u = 0; u < 4; ++u) pr[u] = j + u < info.z ? __ldg(p.enc_pair + j + u) : make_uint2(0u, 0u);
#pragma unroll
                            for (uint32_t u = 0; u < 4; ++u) {
                                const bool kept = pr[u].y >> 31;
                                v0[u] = kept ? __ldcg(rd0 + pr[u].x) : 0u;
                                v1[u] = (kept && two) ? __ldcg(rd1 + pr[u].x) : 0u;
                            }
#pragma unroll
                            for (uint32_t u = 0; u < 4; ++u) {
                                const float c = (float)(pr[u].y & 0x7FFFFFFFu) / nc;
                                if (v0[u]) m0 += fminf((float)v0[u] / nr0, c);
                                if (v1[u]) m1 += fminf((float)v1[u] / nr1, c);
                            }
                        }
                        if (m0 / (l_s[r0].x + lc - m0) < bound) M &= ~(1u << r0);
                        if (two && m1 / (l_s[r1].x + lc - m1) < bound) M &= ~(1u << r1);
                    }
                }
                if (cj_on && M && !(M & (M - 1)) && info.z - info.y == 1) {       // certain join?  (see the kernel comment)
                    const uint32_t r = __ffs(M) - 1;
                    const uint32_t sb = single_s[r];
                    if (sb != kNone && sc_s[r] <= 32767u) {                       // 1024 rows x 32 nnz keep the count below 65536
                        const uint2 pr = __ldg(p.enc_pair + info.y);
                        const uint32_t c = pr.y & 0x7FFFFFFFu;
                        if (pr.x == sb && (pr.y >> 31) && c <= 32u) {
                            cj = true;
                            cj_r = r;
                            cj_cnt = c;
                            M = 0;
                        }
                    }
                }
            }
            s_mask[tid] = M;
            const uint32_t evb = __ballot_sync(0xffffffffu, M != 0);
            const uint32_t lgb = __ballot_sync(0xffffffffu, is_long);
            const uint32_t cjb = __ballot_sync(0xffffffffu, cj);
            if (lane == 0) { s_ev[par_a][wid] = evb; s_long[par_a][wid] = lgb; s_cj[par_a][wid] = cjb; }
            __syncthreads();                                  // S1: verdicts in; the rows stored by earlier steps are ordered before it
            if (tid == 0 && produced != published) st_release_u64(out_ctrl, make_ctrl(s + 1, produced, 0));
            published = produced;
            const uint32_t wmask = s_ev[par_a][lane];
            const uint32_t cjw = s_cj[par_a][lane];
            const uint32_t* longs = s_long[par_a];
            par_a ^= 1;
            uint32_t pre = __popc(wmask), cpre = __popc(cjw);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, pre, d);
                const uint32_t u = __shfl_up_sync(0xffffffffu, cpre, d);
                if ((int)lane >= d) { pre += t; cpre += u; }
            }
            const uint32_t total_ev = __shfl_sync(0xffffffffu, pre, 31);
            const uint32_t total_cj = __shfl_sync(0xffffffffu, cpre, 31);
            // rows [0, limit) of the step are settled: the certain joins among them join, the others move on to the next stage
            auto settle = [&](const uint32_t limit) {
                uint32_t cj_lim = 0;
                if (total_cj) {
                    const uint32_t lw = limit >> 5;
                    cj_lim = limit >= kClusterThreads ? total_cj
                             : __shfl_sync(0xffffffffu, cpre - __popc(cjw), lw) + __popc(__shfl_sync(0xffffffffu, cjw, lw) & ((1u << (limit & 31u)) - 1u));
                }
                uint32_t before = 0;
                if (total_cj) before = __shfl_sync(0xffffffffu, cpre - __popc(cjw), wid) + __popc(__shfl_sync(0xffffffffu, cjw, wid) & ((1u << lane) - 1u));
                if (tid < limit) {
                    if (cj) {
                        atomicAdd(&s_add[cj_r], cj_cnt);
                        p.cluster_ids[pos1] = base + cj_r;
                    } else {
                        out[produced + tid - before] = pos1;
                    }
                }
                produced += limit - cj_lim;
                if (cj_lim) {                                  // (uniform) the representatives' counts; nothing derived from them changes
                    tr_joins += cj_lim;
                    dbg_cj += cj_lim;
                    __syncthreads();
                    if (tid < kStageReps && s_add[tid]) {
                        const uint32_t c2 = sc_s[tid] + s_add[tid];
                        s_add[tid] = 0;
                        sc_s[tid] = c2;
                        repd[(size_t)tid * p.nb + single_s[tid]] = c2;
                        sq_s[tid] = c2 * c2;
                        tot_s[tid] = c2;
                        nr_s[tid] = sqrtf((float)(c2 * c2));
                    }
                    __syncthreads();
                }
            };
            if (total_ev == 0) {
                settle(take);
                cursor += take;
                if (fresh) {
                    __syncthreads();
                    if (tid == 0) st_release_u64(out_ctrl, make_ctrl(s + 1, produced, 0));
                    published = produced;
                }
                dbg_cyc_stream += clock64() - ts0;
                continue;
            }
            auto nth_ev = [&](uint32_t rk) -> uint32_t {
                const uint32_t w = __ffs(__ballot_sync(0xffffffffu, pre > rk)) - 1;
                const uint32_t before = __shfl_sync(0xffffffffu, pre - __popc(wmask), w);
                const uint32_t bits = __shfl_sync(0xffffffffu, wmask, w);
                return w * 32 + __fns(bits, 0, (int)(rk - before) + 1);
            };
            const uint32_t evaluated = min(total_ev, kWarps * kCpwB);
            const uint32_t take_eff = total_ev > evaluated ? nth_ev(evaluated) : take;
            tr_events += evaluated;
            // (measured dead end: fetching position -> pos_info of a warp's four event rows together and the next row's runs while the
            //  current one is decided -- 16 dependent round trips down to 7 -- was 5 % SLOWER: the extra live state spills at 64 registers)
            const long long te0 = clock64();
#pragma unroll 1
            for (uint32_t q = 0; q < kCpwB; ++q) {
                const uint32_t rk = q * kWarps + wid;
                if (rk < evaluated) {                          // warp-uniform
                    const uint32_t k = nth_ev(rk);
                    const uint4 kinfo = __ldg(p.pos_info + __ldcg(in + cursor + k));
                    uint32_t mk = s_mask[k];
                    const uint32_t kn = kinfo.z - kinfo.y;
                    // (a warp walks 32 runs per round trip: beyond a few hundred runs the CTA-wide scratch is faster, bound included)
                    // (measured: sending every row of > 256 runs to the CTA-wide scratch instead costs 2x -- 88 M scratch passes on the
                    //  2^20-row graph where the warps reject all but 3 M -- so the warps keep the bound; kept as a switch)
                    const bool unmasked = false;
                    if (!unmasked && ((longs[k >> 5] >> (k & 31)) & 1u)) mk = warp_shared_mask(kinfo, mk, nullptr);
                    if (kn > 64) {
                        // the sparse similarity settles most long rows here, a warp each; the CTA-wide scratch pass below is for
                        // the pairs inside alpha +- 1e-3
                        // (one pass over the row per candidate: with many candidates on a very long row -- the hub rows against
                        //  each other at the end of the order, 2 ms per row -- the scratch pass, all candidates at once, is the cheaper one)
                        uint32_t cand = mk;
                        // (measured on the 2^20-row graph: threshold 24 -> 4.7 s, no threshold -> 3.6 s: the scratch pass serialises the CTA)
                        bool exact_needed = unmasked || __popc(mk) * ((kn + 127u) >> 7) > 192u;
                        if (mk && !exact_needed) fast_decide_long(kinfo, mk, cand, exact_needed);
                        if (cand && !exact_needed) {           // its highest bit is a certain acceptance, everything below a rejection
                            if (lane == 0) atomicMin(&s_first[par_b], (k << 5) | (31u - __clz(cand)));
                            cand = 0;
                        }
                        if (lane == 0) {
                            s_qmask[rk] = cand;
                            s_qun[rk] = unmasked;
                            if (cand) s_anyq = 1;
                        }
                        continue;
                    }
                    if (lane == 0) s_qmask[rk] = 0;
                    uint2* wr = s_runs64 + wid * 64;
                    if (mk) {                                  // the row's runs, staged once for this warp
                        __syncwarp();
                        if (lane < kn) wr[lane] = __ldg(p.enc_pair + kinfo.y + lane);
                        if (lane + 32 < kn) wr[lane + 32] = __ldg(p.enc_pair + kinfo.y + lane + 32);
                        __syncwarp();
                        uint32_t acc, amb;
                        fast_decide(kinfo, mk, wr, acc, amb);
                        mk = acc | amb;
                        while (mk) {
                            const uint32_t r = __ffs(mk) - 1;
                            mk &= mk - 1;
                            if (((acc >> r) & 1u) || (kn <= 32 ? evaluate_pair(kinfo, r, wr) : evaluate_medium(kinfo, r, wr, s_wbits + wid * 32))) {
                                if (lane == 0) atomicMin(&s_first[par_b], (k << 5) | r);
                                break;
                            }
                        }
                    }
                }
            }
            __syncthreads();                                  // S2: every short event row decided, the long ones queued
            const long long te1 = clock64();
            // the long event rows, one at a time in list order: expanded once, evaluated against their candidates by a warp each
            const uint32_t anyq = s_anyq;                       // (uniform; mostly 0: the loop below cost 14 % of the 2^22-row run doing nothing)
#pragma unroll 1
            for (uint32_t rk = 0; rk < (anyq ? evaluated : 0u); ++rk) {
                const uint32_t mk = s_qmask[rk];
                if (mk == 0) continue;                         // uniform
                const uint32_t k = nth_ev(rk);
                const uint32_t fcur = s_first[par_b];
                if (fcur != kNone && (fcur >> 5) < k) break;   // an earlier row joins: this one is redone anyway
                const uint4 kinfo = __ldg(p.pos_info + __ldcg(in + cursor + k));
                expand_row(kinfo);
                __syncthreads();
                const uint32_t mk2 = s_qun[rk] ? shared_mask_scratch(kinfo, mk) : mk;      // (uniform)
                if (mk2) scratch_pass(kinfo, mk2, k << 5, &s_first[par_b]);
                clear_row(kinfo);
                __syncthreads();
                ++tr_long;
            }
            if (anyq && tid == 0) s_anyq = 0;                  // (next written after the next step's S1)
            tr_event_cyc += te1 - te0;
            tr_long_cyc += clock64() - te1;
            const uint32_t fv = s_first[par_b];
            par_b ^= 1;
            if (tid == 0) s_first[par_b] = kNone;
            const uint32_t fj = fv == kNone ? kNone : fv >> 5;
            const uint32_t n_rej = fj == kNone ? take_eff : fj;
            settle(n_rej);
            if (fresh && n_rej) {
                __syncthreads();
                if (tid == 0) st_release_u64(out_ctrl, make_ctrl(s + 1, produced, 0));
                published = produced;
            }
            if (fj == kNone) {
                cursor += take_eff;
            } else {
                const uint32_t jr = fv & 31u;
                if (tid == fj) p.cluster_ids[pos1] = base + jr;
                join_absorb_refresh(__ldg(p.pos_info + __ldcg(in + cursor + fj)), jr, kStageReps, nullptr);
                cursor += fj + 1;
                ++tr_joins;
                careful = true;                                // joins come in bursts: go on in rounds of 32 rows
                calm_rows = 0;
                ++dbg_rj_stream;
            }
            dbg_cyc_stream += clock64() - ts0;
        }
        __syncthreads();
        if (tid == 0 && p.trace_ts && s <= p.M) p.trace_ts[3 * s + 2] = now_ns();
        if (tid == 0 && p.trace_stage && s <= 16384) {
            unsigned long long* d = p.trace_stage + 8 * (size_t)s;
            d[0] = dbg_rounds; d[1] = dbg_rj_careful; d[2] = dbg_rj_stream; d[3] = dbg_cj; d[4] = dbg_steps; d[5] = dbg_solo;
            d[6] = (unsigned long long)dbg_cyc_careful; d[7] = (unsigned long long)dbg_cyc_stream;
        }
        if (tid == 0) {
            if (produced == 0) {                 // nothing left for a next stage: this one holds the last cluster
                p.status[0] = base + nrep - 1;
                __threadfence();
                atomicExch(p.status + 1, 1u);
            }
            st_release_u64(out_ctrl, make_ctrl(s + 1, produced, 1));
        }
    }
}

__global__ void init_cluster_state_kernel(uint32_t M, uint32_t zero_rows, uint32_t num_slots, uint32_t* cluster_ids, uint32_t* list1,
                                          unsigned long long* ctrl, uint32_t* status) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
        cluster_ids[i] = i < zero_rows ? 0u : kInf;
        if (i >= zero_rows) list1[i - zero_rows] = (uint32_t)i;      // list of cluster 1: every non-empty row, in order
    }
    if (blockIdx.x == 0) {
        // list ids are >= 1, so id 0 marks "never written"; list 1 is complete from the start
        for (uint32_t sidx = threadIdx.x; sidx < num_slots; sidx += blockDim.x)
            ctrl[sidx] = sidx == (1 % num_slots) ? (((unsigned long long)1 << 33) | ((unsigned long long)(M - zero_rows) << 1) | 1ull) : 0ull;
        if (threadIdx.x < 4) status[threadIdx.x] = 0;
    }
}

int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) ++b;
    return b;
}

// src/rowReordering.cu:911-920
uint32_t clustering_blockdim(uint32_t nb) {
    if (nb < 32) return 32;
    int cand = 32 * static_cast<int>(std::ceil(static_cast<float>(nb / 4) / 32.0f));
    if (cand < 32) cand = 32;
    return cand > 1024 ? 1024u : static_cast<uint32_t>(cand);
}

}  // namespace

int row_reorder(bsmr_plan* plan, float alpha, uint32_t block_size, uint32_t flags) {
    bsmr_ctx* ctx = plan->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t M = plan->M, N = plan->N, nnz = plan->nnz;
    const int sm = ctx->sm_count;
    Workspace* ws = &ctx->ws;
    ws->reset();
    cudaEvent_t e0, e1;
    BSMR_CUDA_OK(cudaEventCreate(&e0));
    BSMR_CUDA_OK(cudaEventCreate(&e1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{e0, e1};
    BSMR_CUDA_OK(cudaEventRecord(e0, st));

    plan->h_dispersions.clear();
    plan->h_cluster_ids.clear();
    std::vector<uint32_t> h_ro(static_cast<size_t>(M) + 1);
    BSMR_CUDA_OK(cudaMemcpyAsync(h_ro.data(), plan->row_offsets.ptr, h_ro.size() * 4, cudaMemcpyDeviceToHost, st));
    BSMR_CUDA_OK(cudaStreamSynchronize(st));

    std::vector<uint32_t>& perm = plan->h_reordered_rows;
    perm.clear();
    if ((flags & 3u) == BSMR_ROW_IDENTITY) {
        // noReorderRow (src/rowReordering.cu:15-46): original order, empty rows removed
        for (uint32_t r = 0; r < M; ++r)
            if (h_ro[r + 1] != h_ro[r]) perm.push_back(r);
        plan->num_clusters = plan->num_clusters_true = 1;
        plan->block_size = 0;
    } else {
        if (block_size == 0) BSMR_TRY(bsmr_calculate_block_size(ctx, M, N, 0, &block_size));
        plan->block_size = block_size;
        const uint32_t nb = static_cast<uint32_t>(std::ceil(static_cast<float>(N) / static_cast<float>(block_size)));  // :1035
        const uint32_t bd = clustering_blockdim(nb);
        const uint32_t nw = bd / 32;
        const bool exact = (flags & 3u) == BSMR_ROW_EXACT_REDUCE;
        uint32_t first_stride = nw / 2;  // blockDim.x / (2 * warpSize)   (include/cudaUtil.cuh:37)
        if (exact) {
            uint32_t p2 = 1;
            while (p2 < nw) p2 <<= 1;
            first_stride = p2 / 2;
        }
        // which reference warps survive the tree (every warp at most once)
        uint32_t kept_mask = 0;
        {
            std::vector<uint32_t> contrib(64, 0);
            for (uint32_t w = 0; w < 32; ++w) contrib[w] = w < nw ? (1u << w) : 0u;
            for (uint32_t s = first_stride; s >= 1; s >>= 1)
                for (uint32_t w = 0; w < s; ++w) contrib[w] |= contrib[w + s];
            kept_mask = contrib[0];
        }
        // per-warp dense scratch for rows with more than 32 column blocks, when it fits next to the representative
        const uint32_t scratch_entries = (nb + 7u) & ~7u;
        const size_t base_smem = (static_cast<size_t>(nb) * 2 + 1024 + 32) * 4;
        const bool use_scratch = block_size <= 65535u && base_smem + static_cast<size_t>(scratch_entries) * 2 * 32 <= 200 * 1024;
        const size_t cluster_smem = base_smem + (use_scratch ? static_cast<size_t>(scratch_entries) * 2 * 32 : 0);
        if (cluster_smem > 200 * 1024) {
            set_error("row reorder: %u column blocks need %zu bytes of shared memory; raise block_size", nb, cluster_smem);
            return BSMR_ERR_UNSUPPORTED;
        }

        TmpBuf<uint8_t> temp(ws);
        auto ensure_temp = [&](size_t bytes) -> int {
            if (bytes > temp.count) return temp.alloc(bytes + bytes / 8 + 256);
            return BSMR_OK;
        };
        // ---- sparse encodings ----
        TmpBuf<uint64_t> keys_a(ws), keys_b(ws), ukeys(ws);
        TmpBuf<uint32_t> counts(ws), num_runs_d(ws), runs_per_row(ws), enc_ptr(ws), disp(ws), row_sq(ws), row_tot(ws), enc_blk(ws);
        TmpBuf<uint2> enc_pair(ws);
        BSMR_TRY(runs_per_row.alloc(static_cast<size_t>(M) + 1));
        BSMR_TRY(enc_ptr.alloc(static_cast<size_t>(M) + 1));
        BSMR_TRY(disp.alloc(M ? M : 1));
        BSMR_TRY(row_sq.alloc(M ? M : 1));
        BSMR_TRY(row_tot.alloc(M ? M : 1));
        BSMR_CUDA_OK(cudaMemsetAsync(runs_per_row.ptr, 0, runs_per_row.bytes(), st));
        uint32_t num_runs = 0;
        if (nnz) {
            BSMR_TRY(keys_a.alloc(nnz)); BSMR_TRY(keys_b.alloc(nnz)); BSMR_TRY(ukeys.alloc(nnz)); BSMR_TRY(counts.alloc(nnz));
            BSMR_TRY(num_runs_d.alloc(1));
            enc_keys_kernel<<<grid_for((uint64_t)M * 32, kThreads, sm), kThreads, 0, st>>>(M, plan->row_offsets.ptr, plan->col_indices.ptr, block_size, keys_a.ptr);
            ctx->launches++;
            // rows are already grouped; sort the block ids inside the rows (columns of a row need not be sorted,
            // the .mtx loader keeps file order: src/Matrix.cpp:467-470)
            cub::DoubleBuffer<uint64_t> dk(keys_a.ptr, keys_b.ptr);
            size_t tb = 0;
            const int end_bit = 32 + bits_for(M ? M - 1 : 0);
            BSMR_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, tb, dk, static_cast<int64_t>(nnz), 0, end_bit, st));
            BSMR_TRY(ensure_temp(tb));
            BSMR_CUDA_OK(cub::DeviceRadixSort::SortKeys(temp.ptr, tb, dk, static_cast<int64_t>(nnz), 0, end_bit, st));
            ctx->launches++;
            BSMR_CUDA_OK(cub::DeviceRunLengthEncode::Encode(nullptr, tb, dk.Current(), ukeys.ptr, counts.ptr, num_runs_d.ptr, static_cast<int64_t>(nnz), st));
            BSMR_TRY(ensure_temp(tb));
            BSMR_CUDA_OK(cub::DeviceRunLengthEncode::Encode(temp.ptr, tb, dk.Current(), ukeys.ptr, counts.ptr, num_runs_d.ptr, static_cast<int64_t>(nnz), st));
            ctx->launches++;
            BSMR_CUDA_OK(cudaMemcpyAsync(&num_runs, num_runs_d.ptr, 4, cudaMemcpyDeviceToHost, st));
            BSMR_CUDA_OK(cudaStreamSynchronize(st));
            BSMR_TRY(enc_blk.alloc(num_runs));
            BSMR_TRY(enc_pair.alloc(num_runs));
            runs_split_kernel<<<grid_for(num_runs, kThreads, sm), kThreads, 0, st>>>(ukeys.ptr, counts.ptr, num_runs, bd, kept_mask, enc_blk.ptr, enc_pair.ptr,
                                                                                runs_per_row.ptr);
            ctx->launches++;
        }
        {
            size_t tb = 0;
            BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, runs_per_row.ptr, enc_ptr.ptr, static_cast<size_t>(M) + 1, st));
            BSMR_TRY(ensure_temp(tb));
            BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, runs_per_row.ptr, enc_ptr.ptr, static_cast<size_t>(M) + 1, st));
            ctx->launches++;
        }
        TmpBuf<uint32_t> asc_a(ws), asc_b(ws), dk_a(ws), dk_b(ws), zero_cnt(ws);
        BSMR_TRY(asc_a.alloc(M ? M : 1)); BSMR_TRY(asc_b.alloc(M ? M : 1)); BSMR_TRY(dk_a.alloc(M ? M : 1)); BSMR_TRY(dk_b.alloc(M ? M : 1));
        BSMR_TRY(zero_cnt.alloc(1));
        BSMR_CUDA_OK(cudaMemsetAsync(zero_cnt.ptr, 0, 4, st));
        const uint32_t* asc = asc_a.ptr;
        uint32_t zero_rows = 0;
        if (M) {
            dispersion_kernel<<<grid_for((uint64_t)M * 32, kThreads, sm), kThreads, 0, st>>>(M, plan->row_offsets.ptr, enc_ptr.ptr, enc_blk.ptr, counts.ptr,
                                                                             block_size, bd, kept_mask, disp.ptr, row_sq.ptr, row_tot.ptr);
            // stable ascending sort of the rows by dispersion (:1055-1062)
            BSMR_CUDA_OK(cudaMemcpyAsync(dk_a.ptr, disp.ptr, static_cast<size_t>(M) * 4, cudaMemcpyDeviceToDevice, st));
            iota_kernel<<<grid_for(M, kThreads, sm), kThreads, 0, st>>>(asc_a.ptr, M);
            cub::DoubleBuffer<uint32_t> k2(dk_a.ptr, dk_b.ptr), v2(asc_a.ptr, asc_b.ptr);
            size_t tb = 0;
            BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k2, v2, static_cast<int64_t>(M), 0, 32, st));
            BSMR_TRY(ensure_temp(tb));
            BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp.ptr, tb, k2, v2, static_cast<int64_t>(M), 0, 32, st));
            asc = v2.Current();
            // rows with dispersion 0 are exactly the empty rows and sort first (:939-949)
            count_zero_kernel<<<grid_for(M, kThreads, sm), kThreads, 0, st>>>(disp.ptr, M, zero_cnt.ptr);
            ctx->launches += 4;
            BSMR_CUDA_OK(cudaMemcpyAsync(&zero_rows, zero_cnt.ptr, 4, cudaMemcpyDeviceToHost, st));
            BSMR_CUDA_OK(cudaStreamSynchronize(st));
        }

        // ---- clustering ----
        TmpBuf<uint32_t> cluster_ids(ws), lists(ws), status(ws);
        TmpBuf<unsigned long long> ctrl(ws);
        TmpBuf<uint4> pos_info(ws);
        TmpBuf<unsigned long long> trace(ws);
#ifdef BSMR_DEBUG
        const bool want_trace = std::getenv("BSMR_TRACE") != nullptr;     // probe builds (make DEBUG=1): per-cluster time stamps and counters
#else
        const bool want_trace = false;
#endif
        BSMR_TRY(trace.alloc(12));
        BSMR_CUDA_OK(cudaMemsetAsync(trace.ptr, 0, trace.bytes(), st));
        TmpBuf<unsigned long long> trace_ts(ws);
        if (want_trace) {
            BSMR_TRY(trace_ts.alloc(3 * (static_cast<size_t>(M) + 2)));
            BSMR_CUDA_OK(cudaMemsetAsync(trace_ts.ptr, 0, trace_ts.bytes(), st));
        }
        BSMR_TRY(cluster_ids.alloc(M ? M : 1));
        BSMR_TRY(status.alloc(4));
        uint32_t clusters_true = 0;
        // Which kernel: graph-shaped inputs (>= 2^15 non-empty rows of at most 128 nnz on average) run the stage kernel -- 32 consecutive
        // clusters per CTA -- everything else the cluster-per-CTA kernel.  BSMR_ROW_STAGE_ON / _OFF force either (same permutation).
        const size_t stage_smem = (static_cast<size_t>((nb + 3u) & ~3u) + kStageReps * 1024 + kStageReps * 32 + 1024) * 4 + static_cast<size_t>((nb + 3u) & ~3u) * 2;
        // (dynamic + the kernel's static shared memory against the opt-in limit: a kernel that does not fit must fall back, not fail)
        cudaFuncAttributes stage_attr{};
        BSMR_CUDA_OK(cudaFuncGetAttributes(&stage_attr, bsa_stage_kernel));
        int smem_optin = 0;
        BSMR_CUDA_OK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        const bool stage_fits = stage_smem + stage_attr.sharedSizeBytes <= static_cast<size_t>(smem_optin) && alpha >= 0.0f && block_size <= 65535u &&
                                nb <= kStageTerms * bd;
        // (not "the per-warp scratch does not fit": at 2^23 rows the block size is 3314 and it does -- the choice is about the rows)
        const bool graph_sized = (M - zero_rows) >= (1u << 15) && static_cast<uint64_t>(nnz) <= 128ull * (M - zero_rows);
        const bool use_stage = stage_fits && ((flags & BSMR_ROW_STAGE_ON) ? true : (flags & BSMR_ROW_STAGE_OFF) ? false
                                              : (graph_sized && !(flags & (BSMR_ROW_THREAD_PRUNE_ON | BSMR_ROW_THREAD_PRUNE_OFF))));
        void* const kernel_fn = use_stage ? reinterpret_cast<void*>(bsa_stage_kernel) : reinterpret_cast<void*>(bsa_cluster_kernel);
        const size_t smem = use_stage ? stage_smem : cluster_smem;
        TmpBuf<uint32_t> repd(ws);
        if (M > zero_rows) {
            BSMR_CUDA_OK(cudaFuncSetAttribute(kernel_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            int per_sm = 0;
            BSMR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel_fn, kClusterThreads, smem));
            if (per_sm < 1) {
                set_error("row reorder: clustering kernel does not fit on an SM (smem %zu)", smem);
                return BSMR_ERR_UNSUPPORTED;
            }
            // one CTA per live cluster; never more CTAs than rows
            int grid = sm * per_sm;
            if (static_cast<uint32_t>(grid) > M - zero_rows) grid = static_cast<int>(M - zero_rows);
            const uint32_t num_slots = static_cast<uint32_t>(grid) + 1;
            const uint32_t list_cap = M - zero_rows;
            BSMR_TRY(lists.alloc(static_cast<size_t>(num_slots) * list_cap));
            BSMR_TRY(ctrl.alloc(num_slots));
            init_cluster_state_kernel<<<grid_for(M, kThreads, sm), kThreads, 0, st>>>(
                M, zero_rows, num_slots, cluster_ids.ptr, lists.ptr + static_cast<size_t>(1 % num_slots) * list_cap, ctrl.ptr, status.ptr);
            BSMR_TRY(pos_info.alloc(M));
            pos_info_kernel<<<grid_for(M, kThreads, sm), kThreads, 0, st>>>(M, asc, enc_ptr.ptr, row_sq.ptr, row_tot.ptr, pos_info.ptr);
            ctx->launches++;
            ClusterParams cp{};
            cp.trace = want_trace ? trace.ptr : nullptr;
            cp.trace_ts = want_trace ? trace_ts.ptr : nullptr;
            TmpBuf<unsigned long long> trace_stage(ws);
            if (want_trace) {
                BSMR_TRY(trace_stage.alloc(8 * 16385));
                BSMR_CUDA_OK(cudaMemsetAsync(trace_stage.ptr, 0, trace_stage.bytes(), st));
                cp.trace_stage = trace_stage.ptr;
            }
            cp.bd_mask = (bd & (bd - 1)) == 0 ? bd - 1 : 0u;
            cp.kept_mask = kept_mask; cp.pos_info = pos_info.ptr; cp.scratch = use_scratch ? scratch_entries : 0u;
            cp.M = M; cp.nb = nb; cp.bd = bd; cp.first_stride = first_stride; cp.zero_rows = zero_rows; cp.alpha = alpha;
            cp.list_cap = list_cap; cp.num_slots = num_slots;
            cp.asc = asc; cp.enc_ptr = enc_ptr.ptr; cp.enc_blk = enc_blk.ptr; cp.counts = counts.ptr; cp.row_sq = row_sq.ptr;
            cp.enc_pair = enc_pair.ptr;
            // the thread-prune step pays where most pairs are rejected by the bounds (graphs: many column blocks, short rows);
            // where rows are long (nips: 777 blocks each) every candidate needs a warp anyway and the scratch path is faster
            // Measured (R-MAT, profiles/r02e_* against r02l_*): 2^20 rows 61.4 -> 38.8 s, 2^19 rows 16.0 -> 14.2 s, but 2^17 rows
            // 1.8 -> 3.6 s (three barriers per step instead of two, and the pipeline is latency-bound there): on from 2^18 rows.
            const bool prune_sized = !use_scratch && (M - zero_rows) >= (1u << 18);
            cp.thread_prune = (flags & BSMR_ROW_THREAD_PRUNE_ON) ? 1u : (flags & BSMR_ROW_THREAD_PRUNE_OFF) ? 0u : (prune_sized ? 1u : 0u);
            if (use_stage) {
                BSMR_TRY(repd.alloc(static_cast<size_t>(grid) * kStageReps * nb));
                cp.repd = repd.ptr;
            }
            cp.cluster_ids = cluster_ids.ptr; cp.lists = lists.ptr; cp.ctrl = ctrl.ptr; cp.status = status.ptr;
            void* args[] = {&cp};
            BSMR_CUDA_OK(cudaEventRecord(ctx->ev0, st));
            BSMR_CUDA_OK(cudaLaunchCooperativeKernel(kernel_fn, dim3(grid), dim3(kClusterThreads), args, smem, st));
            BSMR_CUDA_OK(cudaEventRecord(ctx->ev1, st));
            ctx->launches += 2;
            uint32_t h_status[4] = {0, 0, 0, 0};
            BSMR_CUDA_OK(cudaMemcpyAsync(h_status, status.ptr, sizeof(h_status), cudaMemcpyDeviceToHost, st));
            BSMR_CUDA_OK(cudaStreamSynchronize(st));
            if (h_status[2] != 0 || h_status[1] == 0) {
                set_error("row reorder: clustering pipeline stalled (watchdog %u, finished %u)", h_status[2], h_status[1]);
                return BSMR_ERR_CUDA;
            }
            clusters_true = h_status[0];
            if (want_trace) {
                unsigned long long h_tr[12];
                BSMR_CUDA_OK(cudaMemcpy(h_tr, trace.ptr, sizeof(h_tr), cudaMemcpyDeviceToHost));
                if (use_stage)
                    fprintf(stderr, "[bsmr trace] stage kernel: steps %llu rows-in-steps %llu joins %llu polls %llu event rows %llu founded %llu long rows through the scratch %llu | "
                                    "Mcycles (thread 0 of every CTA): busy %.1f poll %.1f founding steps %.1f event phase A %.1f scratch pass %.1f\n",
                            h_tr[0], h_tr[1], h_tr[2], h_tr[3], h_tr[4], h_tr[5], h_tr[6], h_tr[7] / 1e6, h_tr[8] / 1e6, h_tr[9] / 1e6, h_tr[10] / 1e6, h_tr[11] / 1e6);
                else
                fprintf(stderr, "[bsmr trace] clustering: stage kernel %d (then: Mcycles poll = events evaluated, eval = reps founded) rows %u nb %u bd %u grid %d scratch %d | clusters %u steps %llu candidates %llu joins %llu "
                                "polls %llu size-bound rejects seen by warp 0 of every CTA %llu | Mcycles: poll %.1f eval %.1f update %.1f busy(all CTAs) %.1f\n",
                        (int)use_stage, M, nb, bd, grid, (int)use_scratch, clusters_true, h_tr[0], h_tr[1], h_tr[2], h_tr[3], h_tr[8], h_tr[4] / 1e6, h_tr[5] / 1e6,
                        h_tr[6] / 1e6, h_tr[7] / 1e6);
                const uint32_t nc = use_stage ? (clusters_true + kStageReps - 1) / kStageReps : (clusters_true < M ? clusters_true : M);
                std::vector<unsigned long long> ts(3 * (static_cast<size_t>(nc) + 1));
                BSMR_CUDA_OK(cudaMemcpy(ts.data(), trace_ts.ptr, ts.size() * 8, cudaMemcpyDeviceToHost));
                if (nc >= 2) {
                    const unsigned long long t0 = ts[3];
                    auto us = [&](unsigned long long t) { return (double)(t - t0) / 1e3; };
                    fprintf(stderr, "[bsmr trace]   cluster: start / first publish / end (us)\n");
                    if (use_stage) {
                        std::vector<unsigned long long> sd(8 * 16385);
                        BSMR_CUDA_OK(cudaMemcpy(sd.data(), trace_stage.ptr, sd.size() * 8, cudaMemcpyDeviceToHost));
                        fprintf(stderr, "[bsmr stage] stage: careful rounds / joins in careful / joins in streaming / certain joins / streaming steps / solo rows / Mcyc careful / Mcyc streaming\n");
                        for (uint32_t c = 1; c <= nc && c <= 16384; c = c < 8 ? c + 1 : (c < 400 ? c + 8 : c + 500))
                            fprintf(stderr, "[bsmr stage] %6u: %llu %llu %llu %llu %llu %llu %.1f %.1f\n", c, sd[8 * c], sd[8 * c + 1], sd[8 * c + 2], sd[8 * c + 3],
                                    sd[8 * c + 4], sd[8 * c + 5], sd[8 * c + 6] / 1e6, sd[8 * c + 7] / 1e6);
                    }
#ifdef BSMR_DEBUG
                    const char* tv = std::getenv("BSMR_TRACE");
                    const bool all = tv && tv[0] == 'a';
#else
                    const bool all = false;                    // (the release library reads no environment variable)
#endif
                    for (uint32_t c = 1; c <= nc; c = (all || c < 8) ? c + 1 : c + (nc / 12 ? nc / 12 : 1))
                        fprintf(stderr, "[bsmr trace]   %6u: %10.1f %10.1f %10.1f\n", c, us(ts[3 * c]), us(ts[3 * c + 1]), us(ts[3 * c + 2]));
                    fprintf(stderr, "[bsmr trace]   %6u: %10.1f %10.1f %10.1f\n", nc, us(ts[3 * nc]), us(ts[3 * nc + 1]), us(ts[3 * nc + 2]));
                }
            }
            BSMR_CUDA_OK(cudaEventElapsedTime(&plan->cluster_ms, ctx->ev0, ctx->ev1));
        } else if (M) {
            init_cluster_state_kernel<<<grid_for(M, kThreads, sm), kThreads, 0, st>>>(M, zero_rows, 0, cluster_ids.ptr, nullptr, nullptr, status.ptr);
            ctx->launches++;
        }

        // ---- permutation = stable sort of the positions by cluster id (:986-995) ----
        TmpBuf<uint32_t> ck_b(ws), idx_a(ws), idx_b(ws), perm_d(ws);
        std::vector<uint32_t> h_sorted_ids, h_indices, h_perm;
        if (M) {
            BSMR_TRY(ck_b.alloc(M)); BSMR_TRY(idx_a.alloc(M)); BSMR_TRY(idx_b.alloc(M)); BSMR_TRY(perm_d.alloc(M));
            // keep an unsorted copy of the ids for the diagnostics accessor
            plan->h_cluster_ids.resize(M);
            std::vector<uint32_t> h_ids_by_pos(M), h_asc(M);
            BSMR_CUDA_OK(cudaMemcpyAsync(h_ids_by_pos.data(), cluster_ids.ptr, static_cast<size_t>(M) * 4, cudaMemcpyDeviceToHost, st));
            BSMR_CUDA_OK(cudaMemcpyAsync(h_asc.data(), asc, static_cast<size_t>(M) * 4, cudaMemcpyDeviceToHost, st));
            iota_kernel<<<grid_for(M, kThreads, sm), kThreads, 0, st>>>(idx_a.ptr, M);
            cub::DoubleBuffer<uint32_t> k3(cluster_ids.ptr, ck_b.ptr), v3(idx_a.ptr, idx_b.ptr);
            size_t tb = 0;
            BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k3, v3, static_cast<int64_t>(M), 0, 32, st));
            BSMR_TRY(ensure_temp(tb));
            BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp.ptr, tb, k3, v3, static_cast<int64_t>(M), 0, 32, st));
            gather_kernel<<<grid_for(M, kThreads, sm), kThreads, 0, st>>>(asc, v3.Current(), M, perm_d.ptr);
            ctx->launches += 3;
            h_sorted_ids.resize(M); h_indices.resize(M); h_perm.resize(M);
            BSMR_CUDA_OK(cudaMemcpyAsync(h_sorted_ids.data(), k3.Current(), static_cast<size_t>(M) * 4, cudaMemcpyDeviceToHost, st));
            BSMR_CUDA_OK(cudaMemcpyAsync(h_indices.data(), v3.Current(), static_cast<size_t>(M) * 4, cudaMemcpyDeviceToHost, st));
            BSMR_CUDA_OK(cudaMemcpyAsync(h_perm.data(), perm_d.ptr, static_cast<size_t>(M) * 4, cudaMemcpyDeviceToHost, st));
            plan->h_dispersions.resize(M);
            BSMR_CUDA_OK(cudaMemcpyAsync(plan->h_dispersions.data(), disp.ptr, static_cast<size_t>(M) * 4, cudaMemcpyDeviceToHost, st));
            BSMR_CUDA_OK(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < M; ++i) plan->h_cluster_ids[h_asc[i]] = h_ids_by_pos[i];
            // cluster_cnt = cluster_ids[indices[rows-1]] + (zero_row_idx != 0)  -- the index is applied to the SORTED key
            // array (src/rowReordering.cu:996); kept verbatim as the "reference" cluster count
            plan->num_clusters = static_cast<int>(h_sorted_ids[h_indices[M - 1]]) + (zero_rows != 0 ? 1 : 0);
            plan->num_clusters_true = static_cast<int>(clusters_true);
            // strip the leading empty rows (:1081-1090)
            uint32_t first = 0;
            while (first < M && h_ro[h_perm[first] + 1] - h_ro[h_perm[first]] == 0) ++first;
            perm.assign(h_perm.begin() + first, h_perm.end());
        } else {
            plan->num_clusters = plan->num_clusters_true = 0;
        }
    }

    BSMR_TRY(plan->reordered_rows.alloc(perm.size()));
    if (!perm.empty())
        BSMR_CUDA_OK(cudaMemcpyAsync(plan->reordered_rows.ptr, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice, st));
    // numRowPanels_ = ceil(reorderedRows.size() / ROW_PANEL_SIZE)   (src/BSMR.cpp:48)
    plan->num_row_panels = static_cast<uint32_t>((perm.size() + kPanel - 1) / kPanel);
    BSMR_CUDA_OK(cudaEventRecord(e1, st));
    BSMR_CUDA_OK(cudaEventSynchronize(e1));
    BSMR_CUDA_OK(cudaGetLastError());
    BSMR_CUDA_OK(cudaEventElapsedTime(&plan->row_ms, e0, e1));
    return BSMR_OK;
}

}  // namespace bsmr
