// Sparse-residual SDDMM kernel (CUDA cores) for sm_100a.
//
// Replaces the reference's sddmm_gpu_sparse_block_2_2threadOneData_shuffle and
// sddmm_gpu_sparse_remainder_k32_2threadOneData_shuffle (src/sddmmKernel.cu:1994-2104,
// 2109-2199).  Those use 2 threads per nnz, a 16x36 smem A tile rebuilt every 32 k and
// two 16-byte loads per thread per step.
//
// residual_rows_kernel (the fast path, K in {32, 64, 128, 256, 512}):
//   * the residual entries are stored ROW-sorted (reordered row, then CSR position); a warp
//     owns 32 consecutive entries whose (A row, B column, P index) triples are read with one
//     coalesced load each and handed around by shuffles
//   * a group of LPN lanes (8 / 16 / 32, chosen from K) computes one entry; the A row lives
//     in REGISTERS (K/LPN floats per lane) and is reloaded only when the row changes, so the
//     steady state moves exactly one K-vector of B per nnz through L1 -- the first version
//     re-read the A row per nnz and was L1-bound (ncu: l1tex 76 %, profiles/r01a_*)
//   * every lane loads 128-bit pieces of the B column (a K-vector is one or more perfectly
//     coalesced 128..512-byte requests); UNROLL entries are in flight before the first FMA
//   * the dot products of a whole chunk are reduced together: lane-local partials p[it] for
//     the LPN entries a group handles, then a transposing butterfly (LPN-1 shuffles per LPN
//     results instead of LPN*log2(LPN)); lane s ends up with the finished value of entry
//     s*G + group and stores it
// No shared memory, no block barrier.
// residual_sddmm_kernel / residual_sddmm_generic_kernel: any other K (runtime K loop; A from L1).
//
// Roofline: HBM-bound on compulsory traffic; what it actually stresses is L2 -> SM gather
// bandwidth (K*4 bytes of B per nnz), see DESIGN.md.
#include <cuda_fp16.h>

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace bsmr {
namespace {

constexpr int kResThreads = 256;
constexpr int kWarpsPerCta = kResThreads / 32;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}

// ---- L2 eviction priorities for gathers that do not fit in L2 (HINT variants) ----------------------------------
// On a power-law pattern whose B is far larger than L2 (R-MAT 2^22, K = 256: B = 4.3 GB) the residual kernel runs at
// DRAM speed (ncu profiles/r01h_*: 5.5 TB/s, L2 hit rate 36 %): every K-vector of a cold column that passes through
// L2 pushes out a hub column that would have been hit again.  The plan marks the highest-degree columns whose
// K-vectors fit an L2 budget in a bitmap (hot_columns below); their loads carry evict_last, the A rows, the index
// lists and P (each touched once) carry evict_first / streaming, the cold columns evict_first or the default.
// Measured (profiles/r01h_l2_policy_sweep.md): -5 % at B = 4.3 GB, -7 % at 8.6 GB with evict_first on the cold columns;
// evict_last alone (cold columns on the default priority) is slower than no hints, and at B = 537 MB (70 % L2 hit rate
// as it is) every variant costs 10 % -- so the policy is on only above 2 GiB of B.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg4_hint(const float* p, uint64_t policy) {
    float4 v;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
    return v;
}

// B may be stored as fp16 (bsmr_sddmm_f16b: 11 significant bits like the TF32 operands of the tensor-core kernels, fp32
// products and accumulation): four k of a column are then one 8-byte load, and a K-vector is half the bytes at every
// level of the gather (L1 wavefronts, L2 -> SM, DRAM).
template <typename BT> __device__ __forceinline__ float4 ldb4(const BT* p);
template <> __device__ __forceinline__ float4 ldb4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 ldb4<__half>(const __half* p) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
template <typename BT> __device__ __forceinline__ float4 ldb4_hint(const BT* p, uint64_t policy);
template <> __device__ __forceinline__ float4 ldb4_hint<float>(const float* p, uint64_t policy) { return ldg4_hint(p, policy); }
template <> __device__ __forceinline__ float4 ldb4_hint<__half>(const __half* p, uint64_t policy) {
    uint2 raw;
    asm("ld.global.nc.L2::cache_hint.v2.b32 {%0, %1}, [%2], %3;" : "=r"(raw.x), "=r"(raw.y) : "l"(p), "l"(policy));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Batch (sddmm_gpu_batch, src/sddmmKernel.cu:2764-2848: the reference folds the batch into gridDim.z of its kernels):
// blockIdx.y = batch element; A, B, P advance by whole matrices, the index lists are shared.
struct BatchStride {
    size_t a, b, p;      // elements between consecutive batch elements of A, B, P (0 for a single SDDMM)
};

// LPN: lanes per nnz.  KV: number of float4 pieces per lane (K == LPN*4*KV); KV == 0 -> runtime loop.
template <int LPN, int KV, int UNROLL, typename BT>
__global__ void __launch_bounds__(kResThreads)
residual_sddmm_kernel(const uint32_t K, const float* __restrict__ A, const BT* __restrict__ B,
                      float* __restrict__ P, const uint32_t* __restrict__ res_row,
                      const uint32_t* __restrict__ res_col, const uint32_t* __restrict__ res_out,
                      const uint64_t begin, const uint64_t end, const BatchStride bs) {
    A += blockIdx.y * bs.a;
    B += blockIdx.y * bs.b;
    P += blockIdx.y * bs.p;
    constexpr int G = 32 / LPN;     // entries processed concurrently by one warp
    constexpr int ITERS = 32 / G;   // passes to cover the warp's 32 entries
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sub = lane / LPN;       // which concurrent entry this lane works on
    const uint32_t sl = lane % LPN;        // lane inside the group
    const uint64_t num_chunks = (end - begin + 31) / 32;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarpsPerCta;

    for (uint64_t chunk = warp_global; chunk < num_chunks; chunk += warp_stride) {
        const uint64_t e = begin + chunk * 32 + lane;
        const bool valid = e < end;
        // coalesced metadata loads; invalid lanes point at entry `begin` (always in range)
        const uint64_t es = valid ? e : begin;
        const uint32_t my_row = __ldg(res_row + es);
        const uint32_t my_col = __ldg(res_col + es);
        const uint32_t my_out = res_out ? __ldg(res_out + es) : (uint32_t)es;  // NULL = identity (CSR order)
        float my_res = 0.f;

#pragma unroll 1
        for (int it0 = 0; it0 < ITERS; it0 += UNROLL) {
            float acc[UNROLL];
            if constexpr (KV > 0) {
                float4 av[UNROLL][KV], bv[UNROLL][KV];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int j = (it0 + u) * G + sub;  // entry (lane index) this group computes
                    const uint32_t row = __shfl_sync(0xffffffffu, my_row, j);
                    const uint32_t col = __shfl_sync(0xffffffffu, my_col, j);
                    const float* ap = A + (size_t)row * K + sl * 4;
                    const BT* bp = B + (size_t)col * K + sl * 4;
#pragma unroll
                    for (int v = 0; v < KV; ++v) {
                        av[u][v] = ldg4(ap + v * LPN * 4);
                        bv[u][v] = ldb4<BT>(bp + v * LPN * 4);
                    }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    acc[u] = 0.f;
#pragma unroll
                    for (int v = 0; v < KV; ++v) acc[u] = dot4(av[u][v], bv[u][v], acc[u]);
                }
            } else {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int j = (it0 + u) * G + sub;
                    const uint32_t row = __shfl_sync(0xffffffffu, my_row, j);
                    const uint32_t col = __shfl_sync(0xffffffffu, my_col, j);
                    const float* ap = A + (size_t)row * K;
                    const BT* bp = B + (size_t)col * K;
                    float s = 0.f;
                    for (uint32_t k = sl * 4; k < K; k += LPN * 4) s = dot4(ldg4(ap + k), ldb4<BT>(bp + k), s);
                    acc[u] = s;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float s = acc[u];
#pragma unroll
                for (int w = LPN / 2; w >= 1; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
                // entry j = (it0+u)*G + sub now lives in every lane of group `sub`; hand it to lane j
                const float v = __shfl_sync(0xffffffffu, s, (lane % G) * LPN);
                if ((int)(lane / G) == it0 + u) my_res = v;
            }
        }
        if (valid) P[my_out] = my_res;
    }
}

// Row-sorted fast path.  LPN lanes per entry, KV float4 pieces per lane: K == LPN * 4 * KV.
// HINT: col_hot is the bitmap of the columns to keep in L2 (bit c of word c / 32), cold_first the policy of the others.
#ifndef BSMR_RES_OCC
#define BSMR_RES_OCC 4      // resident CTAs per SM the register budget is set for (64 registers; measured against 3 and 5)
#endif
#ifndef BSMR_RES_UNROLL
#define BSMR_RES_UNROLL 4   // entries in flight per lane group for KV < 4
#endif
#ifndef BSMR_RES_K128_LPN
#define BSMR_RES_K128_LPN 16   // lanes per entry of the fp32 kernel at K = 128: two 16-byte pieces per lane (measured against 32 x one:
                               // 1.20 against 1.35 ms on the 2^20-row graph, 34.8 against 38.9 us on nips)
#endif
#ifndef BSMR_RES_HALVE_LANES
#define BSMR_RES_HALVE_LANES 0  // probe builds: 1 = half the lanes / twice the pieces per lane at the other K as well
#endif

// One 16-byte piece of a B column per lane: 4 fp32 or 8 fp16 values.
template <typename BT> struct BPiece;
template <> struct BPiece<float> {
    static constexpr int EP = 4;                  // elements of K per piece
    using Raw = float4;
    static __device__ __forceinline__ Raw load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ Raw load_hint(const float* p, uint64_t pol) { return ldg4_hint(p, pol); }
    static __device__ __forceinline__ float dot(const float4* a, const Raw& b, float acc) { return dot4(a[0], b, acc); }
};
template <> struct BPiece<__half> {
    static constexpr int EP = 8;
    using Raw = uint4;
    static __device__ __forceinline__ Raw load(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
    static __device__ __forceinline__ Raw load_hint(const __half* p, uint64_t pol) {
        uint4 v;
        asm("ld.global.nc.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
        return v;
    }
    static __device__ __forceinline__ float dot(const float4* a, const Raw& b, float acc) {
        const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
        const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&b.z)), b3 = __half22float2(*reinterpret_cast<const __half2*>(&b.w));
        acc = dot4(a[0], make_float4(b0.x, b0.y, b1.x, b1.y), acc);
        return dot4(a[1], make_float4(b2.x, b2.y, b3.x, b3.y), acc);
    }
};

// Row-sorted fast path.  LPN lanes per entry, KV 16-byte pieces of the B column per lane: K == LPN * EP * KV with EP = 4
// (fp32 B) or 8 (fp16 B) elements per piece.  The kernel is bound by instruction issue as much as by the gather (ncu:
// no unit above 66 %, 13 cycles per entry and SM at K = 128), so the fewer lanes an entry takes the better: at K = 128 a
// 16-lane group (two fp32 pieces, or one fp16 piece per lane) handles two entries per warp instruction.
// HINT: col_hot is the bitmap of the columns to keep in L2 (bit c of word c / 32), cold_first the policy of the others.
template <int LPN, int KV, bool HINT, typename BT>
__global__ void __launch_bounds__(kResThreads, BSMR_RES_OCC)
residual_rows_kernel(const uint32_t K, const float* __restrict__ A, const BT* __restrict__ B,
                     float* __restrict__ P, const uint32_t* __restrict__ res_row,
                     const uint32_t* __restrict__ res_col, const uint32_t* __restrict__ res_out,
                     const uint64_t begin, const uint64_t end, const BatchStride bs,
                     const uint32_t* __restrict__ col_hot, const uint32_t cold_first) {
    A += blockIdx.y * bs.a;
    B += blockIdx.y * bs.b;
    P += blockIdx.y * bs.p;
    using BP = BPiece<BT>;
    constexpr int EP = BP::EP;
    constexpr int AP = EP / 4;       // float4 pieces of the A row per piece of B
    constexpr int G = 32 / LPN;      // entries in flight per warp instruction
    constexpr int H = LPN / 2;
    constexpr int UNROLL0 = KV * AP >= 4 ? 2 : BSMR_RES_UNROLL;
    constexpr int UNROLL = UNROLL0 > 2 * H ? 2 * H : UNROLL0;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sub = lane / LPN;
    const uint32_t sl = lane % LPN;
    const uint64_t num_chunks = (end - begin + 31) / 32;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarpsPerCta;

    uint32_t cur_row = 0xFFFFFFFFu;
    float4 a_cur[KV][AP];
#pragma unroll
    for (int v = 0; v < KV; ++v)
#pragma unroll
        for (int j = 0; j < AP; ++j) a_cur[v][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint64_t pol_hot = 0, pol_cold = 0, pol_stream = 0;
    if constexpr (HINT) {
        pol_hot = l2_policy_evict_last();
        pol_stream = l2_policy_evict_first();
        pol_cold = cold_first ? pol_stream : l2_policy_evict_normal();
    }

    auto load_idx = [&](uint64_t chunk, uint32_t& row, uint32_t& col, uint32_t& out, bool& valid) {
        const uint64_t e = begin + chunk * 32 + lane;
        valid = e < end;
        const uint64_t es = valid ? e : begin;       // idle lanes recompute entry `begin`; never stored
        if constexpr (HINT) {
            row = __ldcs(res_row + es);
            col = __ldcs(res_col + es);
            out = res_out ? __ldcs(res_out + es) : (uint32_t)es;
        } else {
            row = __ldg(res_row + es);
            col = __ldg(res_col + es);
            out = res_out ? __ldg(res_out + es) : (uint32_t)es;   // NULL = identity (CSR order)
        }
    };
    for (uint64_t chunk = warp_global; chunk < num_chunks; chunk += warp_stride) {
        uint32_t my_row, my_col, my_out;
        bool valid;
        load_idx(chunk, my_row, my_col, my_out, valid);     // (fetching them one chunk ahead was measured slower: registers)
        uint32_t my_hot = 0;
        if constexpr (HINT) my_hot = (__ldg(col_hot + (my_col >> 5)) >> (my_col & 31)) & 1u;
        // Pass s (s = 0..LPN-1, entries in list order so that the A row in registers is reused) fills
        // accumulator slot it(s) = (s >> 1) + (s & 1) * H: the two passes of a pair (i, i + H) are adjacent,
        // and the first butterfly step (offset H) is folded in as soon as both are known, which keeps only
        // H accumulators live.
        float q[H];
        const bool upper0 = (sl & H) != 0;
#pragma unroll
        for (int i0 = 0; i0 < H; i0 += UNROLL / 2) {
            typename BP::Raw bv[UNROLL][KV];
            uint32_t rows_u[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int s_pass = 2 * i0 + u;            // list-order pass
                const int j = s_pass * G + sub;           // entry handled by this group in that pass
                rows_u[u] = __shfl_sync(0xffffffffu, my_row, j);
                const uint32_t col = __shfl_sync(0xffffffffu, my_col, j);
                const BT* bp = B + (size_t)col * K + sl * EP;
                if constexpr (HINT) {
                    // the policy operand travels in a uniform register: one priority per warp instruction (with LPN < 32
                    // the G entries of a pass share it: hot if any of them is)
                    const uint32_t hot_j = __shfl_sync(0xffffffffu, my_hot, j);
                    const uint64_t pol = (LPN == 32 ? hot_j != 0 : __any_sync(0xffffffffu, hot_j != 0)) ? pol_hot : pol_cold;
#pragma unroll
                    for (int v = 0; v < KV; ++v) bv[u][v] = BP::load_hint(bp + v * LPN * EP, pol);
                } else {
#pragma unroll
                    for (int v = 0; v < KV; ++v) bv[u][v] = BP::load(bp + v * LPN * EP);
                }
            }
            float d[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (rows_u[u] != cur_row) {              // uniform inside the group (warp-uniform for LPN == 32)
                    const float* ap = A + (size_t)rows_u[u] * K + sl * EP;
#pragma unroll
                    for (int v = 0; v < KV; ++v)
#pragma unroll
                        for (int j = 0; j < AP; ++j)
                            a_cur[v][j] = HINT ? ldg4_hint(ap + v * LPN * EP + 4 * j, pol_stream) : ldg4(ap + v * LPN * EP + 4 * j);
                    cur_row = rows_u[u];
                }
                float acc = 0.f;
#pragma unroll
                for (int v = 0; v < KV; ++v) acc = BP::dot(a_cur[v], bv[u][v], acc);
                d[u] = acc;
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u += 2) {
                const float send = upper0 ? d[u] : d[u + 1];     // slots i (pass 2i) and i + H (pass 2i + 1)
                const float keep = upper0 ? d[u + 1] : d[u];
                q[i0 + u / 2] = keep + __shfl_xor_sync(0xffffffffu, send, H);
            }
            asm volatile("" ::: "memory");   // keep the loads of later batches from being hoisted (register pressure)
        }
        // remaining steps of the transposing butterfly over the LPN lanes of a group: after the step with
        // offset h, lanes whose bit h is clear hold the sums of slots [0,h), the others those of [h,2h)
#pragma unroll
        for (int h = H / 2; h >= 1; h >>= 1) {
            const bool upper = (sl & h) != 0;
#pragma unroll
            for (int i = 0; i < h; ++i) {
                const float send = upper ? q[i] : q[i + h];
                const float keep = upper ? q[i + h] : q[i];
                q[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
            }
        }
        // this lane holds the finished dot product of slot `sl`, i.e. of pass s = 2 * (sl % H) + sl / H
        const int mine = (int)((2 * (sl % H) + sl / H) * G + sub);
        const uint32_t out = __shfl_sync(0xffffffffu, my_out, mine);
        const bool ok = __shfl_sync(0xffffffffu, (int)valid, mine) != 0;
        if (ok) {
            if constexpr (HINT) __stcs(P + out, q[0]);
            else P[out] = q[0];
        }
    }
}

// Any K (no alignment assumption): one lane group of 32, scalar loads.
template <typename BT>
__global__ void __launch_bounds__(kResThreads)
residual_sddmm_generic_kernel(const uint32_t K, const float* __restrict__ A, const BT* __restrict__ B,
                              float* __restrict__ P, const uint32_t* __restrict__ res_row,
                              const uint32_t* __restrict__ res_col, const uint32_t* __restrict__ res_out,
                              const uint64_t begin, const uint64_t end, const BatchStride bs) {
    A += blockIdx.y * bs.a;
    B += blockIdx.y * bs.b;
    P += blockIdx.y * bs.p;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarpsPerCta;
    for (uint64_t e = begin + warp_global; e < end; e += warp_stride) {
        const float* ap = A + (size_t)__ldg(res_row + e) * K;
        const BT* bp = B + (size_t)__ldg(res_col + e) * K;
        float s = 0.f;
        for (uint32_t k = lane; k < K; k += 32) s = fmaf(__ldg(ap + k), static_cast<float>(__ldg(bp + k)), s);
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
        if (lane == 0) P[res_out ? __ldg(res_out + e) : (uint32_t)e] = s;
    }
}

// row_of_nnz[i] = row that owns CSR position i (one warp per row, coalesced writes)
__global__ void expand_rows_kernel(const uint32_t M, const uint32_t* __restrict__ row_offsets,
                                   uint32_t* __restrict__ row_of_nnz) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < M; r += stride) {
        const uint32_t b = __ldg(row_offsets + r), e = __ldg(row_offsets + r + 1);
        for (uint32_t i = b + lane; i < e; i += 32) row_of_nnz[i] = (uint32_t)r;
    }
}

// ---- hub columns (HINT variant of residual_rows_kernel) ----
constexpr int kDegBins = 4096;                // degrees are capped at kDegBins - 1 for the threshold search

__global__ void col_degree_kernel(const uint32_t* __restrict__ col, const uint64_t nnz, uint32_t* __restrict__ deg) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(deg + __ldg(col + i), 1u);
}

__global__ void degree_hist_kernel(const uint32_t* __restrict__ deg, const uint32_t N, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[kDegBins];
    for (int i = threadIdx.x; i < kDegBins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < N; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t d = __ldg(deg + c);
        atomicAdd(&h[d < kDegBins ? d : kDegBins - 1], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kDegBins; i += blockDim.x)
        if (h[i]) atomicAdd(hist + i, h[i]);
}

// bit c of word c / 32 = (degree of column c >= threshold); one warp per word
__global__ void hot_bitmap_kernel(const uint32_t* __restrict__ deg, const uint32_t N, const uint32_t threshold,
                                  uint32_t* __restrict__ bitmap) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t words = ((uint64_t)N + 31) / 32;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w = warp; w < words; w += stride) {
        const uint64_t c = w * 32 + lane;
        const bool hot = c < N && __ldg(deg + c) >= threshold;
        const uint32_t bits = __ballot_sync(0xffffffffu, hot);
        if (lane == 0) bitmap[w] = bits;
    }
}

template <typename BT>
int launch_residual_t(bsmr_ctx* ctx, const ResidualArgs& r) {
    const uint32_t K = r.K;
    const float* dA = r.A;
    const BT* dB = static_cast<const BT*>(r.B);
    const uint64_t begin = r.begin, end = r.end;
    const uint64_t chunks = (end - begin + 31) / 32;
    const uint64_t ctas_needed = (chunks + kWarpsPerCta - 1) / kWarpsPerCta;
    // 8 CTAs of 256 threads = 64 warps = a full SM; grid is a multiple of the SM count
    const uint64_t max_ctas = (uint64_t)ctx->sm_count * 8;
    const bool fast_k = K == 32 || K == 64 || K == 128 || K == 256 || K == 512;
    // residual_rows_kernel: 3 resident CTAs per SM (register budget), persistent grid-stride over the chunks
    const uint64_t cap = fast_k ? (uint64_t)ctx->sm_count * BSMR_RES_OCC : max_ctas;
    const uint32_t batch = r.batch ? r.batch : 1u;
    const BatchStride bs{r.stride_a, r.stride_b, r.stride_p};
    const dim3 grid((unsigned)(ctas_needed < cap ? ctas_needed : cap), batch);
    // 128-bit loads of A, 128-bit (fp32) / 64-bit (fp16) loads of B: every batch element must keep that alignment
    const size_t b_align = 16;
    const bool aligned = (K % (16 / sizeof(BT)) == 0 || !fast_k) && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(dA) % 16 == 0) && (reinterpret_cast<uintptr_t>(dB) % b_align == 0) &&
                         (batch == 1 || ((r.stride_a * sizeof(float)) % 16 == 0 && (r.stride_b * sizeof(BT)) % b_align == 0));
    cudaStream_t st = r.stream;
#define BSMR_ROWS(LPN, KV, HINT) \
    residual_rows_kernel<LPN, KV, HINT, BT><<<grid, kResThreads, 0, st>>>(K, dA, dB, r.P, r.row, r.col, r.out, begin, end, bs, r.col_hot, r.cold_first)
    // lanes per entry x 16-byte pieces per lane: K = LPN * EP * KV (EP = 4 fp32 / 8 fp16 elements per piece).  As few lanes
    // per entry as keeps a piece a full 16 bytes: more entries per warp instruction (the kernel is issue-bound at small K)
    constexpr bool kHalf = sizeof(BT) == 2;
#define BSMR_ROWS_K(HINT)                                                                  \
    do {                                                                                   \
        if constexpr (kHalf) {                                                             \
            if (K == 32) BSMR_ROWS(4, 1, HINT);                                            \
            else if (K == 64) { if (BSMR_RES_HALVE_LANES) BSMR_ROWS(4, 2, HINT); else BSMR_ROWS(8, 1, HINT); } \
            else if (K == 128) BSMR_ROWS(16, 1, HINT);                                     \
            else if (K == 256) BSMR_ROWS(16, 2, HINT);     /* measured: 1.70 against 1.90 ms (32 x 1) on the 2^20-row graph */ \
            else BSMR_ROWS(32, 2, HINT);                                                   \
        } else {                                                                           \
            if (K == 32) { if (BSMR_RES_HALVE_LANES) BSMR_ROWS(4, 2, HINT); else BSMR_ROWS(8, 1, HINT); } \
            else if (K == 64) BSMR_ROWS(8, 2, HINT);       /* measured: mask 98 % 10.6 against 12.4 us hot (16 x 1) */ \
            else if (K == 128) BSMR_ROWS(BSMR_RES_K128_LPN, 128 / (BSMR_RES_K128_LPN * 4), HINT); \
            else if (K == 256) { if (BSMR_RES_HALVE_LANES) BSMR_ROWS(16, 4, HINT); else BSMR_ROWS(32, 2, HINT); } \
            else BSMR_ROWS(32, 4, HINT);                                                   \
        }                                                                                  \
    } while (0)
    if (!aligned) {
        const uint64_t need = (end - begin + kWarpsPerCta - 1) / kWarpsPerCta;
        const dim3 g((unsigned)(need < max_ctas ? need : max_ctas), batch);
        residual_sddmm_generic_kernel<BT><<<g, kResThreads, 0, st>>>(K, dA, dB, r.P, r.row, r.col, r.out, begin, end, bs);
    } else if (fast_k && r.col_hot) {
        // hub columns pinned in L2 (hot_columns below): same kernel, loads and stores carry L2 eviction priorities
        BSMR_ROWS_K(true);
    } else if (fast_k) {
        BSMR_ROWS_K(false);
    } else if (K < 64) {
        residual_sddmm_kernel<8, 0, 2, BT><<<grid, kResThreads, 0, st>>>(K, dA, dB, r.P, r.row, r.col, r.out, begin, end, bs);
    } else if (K < 128) {
        residual_sddmm_kernel<16, 0, 2, BT><<<grid, kResThreads, 0, st>>>(K, dA, dB, r.P, r.row, r.col, r.out, begin, end, bs);
    } else {
        residual_sddmm_kernel<32, 0, 2, BT><<<grid, kResThreads, 0, st>>>(K, dA, dB, r.P, r.row, r.col, r.out, begin, end, bs);
    }
#undef BSMR_ROWS_K
#undef BSMR_ROWS
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

}  // namespace

int launch_residual(bsmr_ctx* ctx, const ResidualArgs& r) {
    if (r.end <= r.begin) return BSMR_OK;
    if (r.K == 0) {
        set_error("K must be positive");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    return r.b_half ? launch_residual_t<__half>(ctx, r) : launch_residual_t<float>(ctx, r);
}

int launch_residual(bsmr_ctx* ctx, uint32_t K, const float* dA, const float* dB, float* dP,
                    const uint32_t* res_row, const uint32_t* res_col, const uint32_t* res_out,
                    uint64_t begin, uint64_t end, const uint32_t* col_hot, uint32_t cold_first) {
    ResidualArgs r{};
    r.K = K; r.A = dA; r.B = dB; r.P = dP; r.row = res_row; r.col = res_col; r.out = res_out;
    r.begin = begin; r.end = end; r.col_hot = col_hot; r.cold_first = cold_first; r.stream = ctx->stream;
    return launch_residual(ctx, r);
}

// fp32 -> fp16 conversion of an operand (round to nearest even), the converter behind bsmr_convert_f32_to_f16
namespace {
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(dst)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < n % 4) dst[n4 * 4 + threadIdx.x] = __float2half_rn(src[n4 * 4 + threadIdx.x]);
}
}  // namespace
int launch_f32_to_f16(bsmr_ctx* ctx, const float* src, void* dst, size_t n, cudaStream_t stream) {
    if (n == 0) return BSMR_OK;
    if ((reinterpret_cast<uintptr_t>(src) % 16) || (reinterpret_cast<uintptr_t>(dst) % 8)) {
        set_error("fp32 -> fp16 conversion needs a 16-byte aligned source and an 8-byte aligned destination");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    const size_t blocks = (n / 4 + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    f32_to_f16_kernel<<<(unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap), 256, 0, stream>>>(src, static_cast<__half*>(dst), n);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

// The columns whose K-vectors the residual kernel asks L2 to keep: the highest-degree columns (degree >= 2) that fit
// plan->l2_hot_budget_mb, as a bitmap; *bitmap = nullptr when B fits the budget anyway (every small workload), when
// the policy is switched off (budget 0) or when K has no fast path.  Built once per (plan, K, budget): a degree count
// over the CSR column indices (once per plan), a capped degree histogram and a threshold chosen on the host.
int hot_columns(bsmr_plan* p, uint32_t K, const uint32_t** bitmap, uint32_t* cold_first, uint32_t b_elem_bytes) {
    *bitmap = nullptr;
    *cold_first = p->l2_cold_first;
    const uint64_t b_bytes = (uint64_t)p->N * K * b_elem_bytes;
    const bool fast_k = K == 32 || K == 64 || K == 128 || K == 256 || K == 512;
    if (!fast_k || p->l2_hot_budget_mb == 0 || p->nnz == 0 || b_bytes <= ((uint64_t)p->l2_hot_min_b_mb << 20)) return BSMR_OK;
    if (p->hot_K == K && p->hot_elem == b_elem_bytes && p->hot_budget_mb == p->l2_hot_budget_mb && p->col_hot.ptr) {
        *bitmap = p->col_hot.ptr;
        return BSMR_OK;
    }
    bsmr_ctx* ctx = p->ctx;
    const int grid = ctx->sm_count * 8;
    if (!p->col_degree.ptr) {
        BSMR_TRY(p->col_degree.alloc(p->N));
        BSMR_CUDA_OK(cudaMemsetAsync(p->col_degree.ptr, 0, p->col_degree.bytes(), ctx->stream));
        col_degree_kernel<<<grid, 256, 0, ctx->stream>>>(p->col_indices.ptr, p->nnz, p->col_degree.ptr);
        ctx->launches++;
    }
    bsmr::DevBuf<uint32_t> d_hist;
    BSMR_TRY(d_hist.alloc(kDegBins));
    BSMR_CUDA_OK(cudaMemsetAsync(d_hist.ptr, 0, d_hist.bytes(), ctx->stream));
    degree_hist_kernel<<<grid, 256, 0, ctx->stream>>>(p->col_degree.ptr, p->N, d_hist.ptr);
    ctx->launches++;
    std::vector<uint32_t> hist(kDegBins);
    BSMR_CUDA_OK(cudaMemcpyAsync(hist.data(), d_hist.ptr, d_hist.bytes(), cudaMemcpyDeviceToHost, ctx->stream));
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    const uint64_t max_cols = ((uint64_t)p->l2_hot_budget_mb << 20) / ((uint64_t)K * b_elem_bytes);
    uint64_t count = 0;
    uint32_t threshold = 0xFFFFFFFFu;        // nothing hot unless a degree class fits the budget
    for (int d = kDegBins - 1; d >= 2; --d) {
        if (count + hist[d] > max_cols) break;
        count += hist[d];
        threshold = (uint32_t)d;
    }
    BSMR_TRY(p->col_hot.alloc(((size_t)p->N + 31) / 32));
    hot_bitmap_kernel<<<grid, 256, 0, ctx->stream>>>(p->col_degree.ptr, p->N, threshold, p->col_hot.ptr);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    p->hot_K = K;
    p->hot_elem = b_elem_bytes;
    p->hot_budget_mb = p->l2_hot_budget_mb;
    p->hot_threshold = threshold;
    p->hot_count = count;
    *bitmap = p->col_hot.ptr;
    return BSMR_OK;
}

int launch_expand_rows(bsmr_ctx* ctx, uint32_t M, uint32_t nnz, const uint32_t* row_offsets, uint32_t* row_of_nnz) {
    if (M == 0 || nnz == 0) return BSMR_OK;
    const uint64_t warps_needed = M;
    const uint64_t ctas = (warps_needed + 7) / 8;
    const uint64_t max_ctas = (uint64_t)ctx->sm_count * 8;
    expand_rows_kernel<<<(int)(ctas < max_ctas ? ctas : max_ctas), 256, 0, ctx->stream>>>(M, row_offsets, row_of_nnz);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

}  // namespace bsmr
