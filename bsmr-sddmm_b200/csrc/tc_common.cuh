// Device-side building blocks shared by the tcgen05 kernels (dense_tc.cu, wide_tc.cu): mbarrier, TMA gather,
// tcgen05 fences / MMA / commit, UMMA shared-memory and instruction descriptors.  Inline PTX only.
#pragma once

#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace bsmr {
namespace tc {

constexpr unsigned long long kWaitLimitNs = 2000000000ull;   // a pipeline wait longer than 2 s is a broken pipeline

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a wrong transaction count must become an error, never a hung GPU.
// kBackoff: after a few immediate polls the warp sleeps between polls.  A warp that spins on try_wait is always "ready"
// and takes issue slots from the warps that share its scheduler (measured in the wide kernel: the epilogue warps ran 3x
// slower next to spinning producers), but every sleep also adds its length to the hand-over latency, so the
// latency-critical hops of a pipeline (converter -> MMA issuer) poll without it.
template <bool kBackoff = true>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t* error_flag, uint32_t code) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        ++spins;
        if (kBackoff && spins > 4) __nanosleep(spins < 64 ? 40 : 200);
        if ((spins & 0xFFFu) == 0) {          // every 4096 polls: wall-clock limit (2 s) on the whole wait
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > kWaitLimitNs) {
                atomicExch(error_flag, code);     // host-mapped word: readable after the trap (common.cuh)
                __threadfence_system();
                __trap();
            }
        }
    }
}

__device__ __forceinline__ void tma_gather4(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int4 rows) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(smem_u32(bar))
        : "memory");
}
// plain tiled load of a box of the tensor map (rows y .. y + box rows - 1, elements x .. x + 31)
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp, SmemDescriptor):
// start address >> 4 in [0,14), LBO = 0, SBO = 1024 bytes (8 rows x 128 B) >> 4 in [32,46),
// version = 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, M x N
// (InstrDescriptor: c_format [4,6) = 1 (F32), a_format [7,10) = 2 (TF32), b_format [10,13) = 2,
//  a_major bit 15 = 0, b_major bit 16 = 0, n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4).
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D fp32 tensor [rows x K] with K contiguous; box = 32 floats of one row; SWIZZLE_128B.
// tf32_rounding: the tensor map carries CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 and the TMA unit rounds every fp32 element to
// TF32 (to nearest) on its way into shared memory -- measured on B200 against cvt.rna.tf32.f32 in a converter pass: same
// maximum error (1.4e-4 at K = 128), mean signed error -3e-7 (tests/tf32_probe.py); with a FLOAT32 map the tensor core
// truncates (mean -6.5e-4).
// box_rows: 1 for tile::gather4 requests (4 arbitrary rows each), 32 / 128 for plain tiled loads of consecutive rows
inline int make_row_gather_map(bsmr_ctx* ctx, const float* base, uint64_t rows, uint64_t K, CUtensorMap* out, bool tf32_rounding = false,
                               uint32_t box_rows = 1) {
    if (!ctx->encode_tiled) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return BSMR_ERR_UNSUPPORTED;
    }
    const cuuint64_t dims[2] = {K, rows};
    const cuuint64_t strides[1] = {K * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const bool tf32_map = tf32_rounding;
    CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
        out, tf32_map ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu K=%llu)", (int)r, (unsigned long long)rows,
                  (unsigned long long)K);
        return BSMR_ERR_CUDA;
    }
    return BSMR_OK;
}


// Predicated 4-byte store P[idx] = bits.  The epilogues store one accumulator per (row, lane) where S has an entry; written
// as `if (...) P[i] = v` the compiler turns each of them into a divergent branch (BSSY / BSYNC per row) as soon as the
// address takes more than an IMAD.WIDE off a kernel parameter -- measured: the mask-form epilogue of the wide kernel went
// from 20 to 35 us per DLMC mask when P became a per-batch-element pointer.  The predicate is explicit here.
__device__ __forceinline__ void st_global_if(float* base, uint32_t idx, uint32_t bits, bool pred) {
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "  setp.ne.b32 p, %2, 0;\n"
        "  @p st.global.b32 [%0], %1;\n"
        "}\n" ::"l"(base + idx), "r"(bits), "r"((uint32_t)pred)
        : "memory");
}

}  // namespace tc
}  // namespace bsmr
