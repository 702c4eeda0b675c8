// Wide row-group SDDMM kernel for sm_100a: a 128-row group of the reordered matrix stays resident in shared memory
// as the M operand of tcgen05.mma (TF32, fp32 accumulators in TMEM), the group's non-empty B columns stream past it
// 256 at a time as the N operand, and the epilogue keeps only the accumulator elements S has: thread = row, one
// 32-bit mask word per 32 columns, the kept values of a (row, tile half) land in ONE contiguous run of P because a
// CSR row is sorted by column and the tile's columns are the group's distinct columns in ascending order.
//
// Why it exists (no counterpart in the reference, whose only tensor-core unit is the 16x16 block of
// src/sddmmKernel.cu:213-351): on matrices that are dense-ish at the scale of a row group (the nips example is 4 %
// dense, DLMC masks 2-30 %) both reference-shaped kernels are bound by the L2 -> SM gather of one K-vector of B per
// nnz (or per 16-row panel column): ~10 TB/s, measured.  Here a B column is fetched once per 128 rows, the gather
// traffic drops by nnz(group) / distinct_columns(group) (8x on nips), and the work that replaces it -- 128 x 256 x K
// MACs per tile whatever the fill -- is what the tensor pipe has to spare.  A group takes this path when
// nnz(group) >= ratio * (256 * tiles + 128) (colreorder.cu: build_wide_format); every other group keeps the BSMR
// dense-block + residual kernels.
//
// Pipeline (13 warps, one CTA per SM, persistent over a contiguous range of tiles):
//   warps 0-7  producers: LDG.128 (8 lanes x 16 B = one 128-byte K-chunk of a column) -> cvt.rna.tf32.f32 (the
//              reference rounds to nearest, wmma::__float_to_tf32; tcgen05 kind::tf32 truncates) -> STS.128 into the
//              SWIZZLE_128B K-major image the MMA expects (16-byte chunk c of row r at chunk c ^ (r & 7)) ->
//              fence.proxy.async -> mbarrier.  Three stages of loads are in flight per lane (register ring).
//              The same warps (re)load the group's A rows, one 16 KB image per 32 floats of K.
//   warp  8    TMEM allocator (512 columns = two 128 x 256 fp32 accumulators) + single-lane MMA issuer
//   warps 9-12 epilogue: warp (quarter q = warp % 4) reads TMEM lanes 32q..32q+31 of the finished accumulator with
//              tcgen05.ld 32x32b.x32, 32 columns at a time, and stores the masked elements.
// Shared memory: K/32 x 16 KB (A, resident) + S x 32 KB (B ring); S is kept small (2-3) on purpose, see launch_wide.
// Roofline: HBM on the compulsory bytes of the step; inside, L2 -> SM traffic (distinct columns x K x 4 per group).
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

#ifndef BSMR_WIDE_NBUF
#define BSMR_WIDE_NBUF 4
#endif

namespace bsmr {
namespace {

using namespace tc;

constexpr int kWRows = BSMR_WIDE_GROUP_ROWS;   // rows of a row group = UMMA M = TMEM lanes
constexpr int kWCols = BSMR_WIDE_TILE_COLS;    // max columns of a wide tile = UMMA N
constexpr int kWChunk = 32;                    // floats of K per stage (128 bytes = one swizzle row)
constexpr int kWAChunkBytes = kWRows * 128;    // 16 KB
constexpr int kWBStageBytes = kWCols * 128;    // 16 KB
constexpr int kWRounds = kWCols / 32;          // LDG.128 per producer lane per stage (a warp instruction covers 4 columns)
constexpr int kWWords = kWCols / 32;           // 32-column chunks of a tile
constexpr int kWAccs = 512 / kWCols;           // TMEM accumulators in rotation
constexpr int kWProducerWarps = 8;
constexpr int kWEpiWarp0 = kWProducerWarps;          // warps 8-11: TMEM lane quarter = warp % 4
constexpr int kWEpiWarps = 4;
constexpr int kWMmaWarp = kWEpiWarp0 + kWEpiWarps;   // warp 12 (13-15 idle: setmaxnreg works on whole warpgroups)
constexpr int kWThreads = 512;
constexpr int kWProducerRegs = 168;                  // 256 threads x 168 + 256 threads x 88 = 64 K registers
constexpr int kWOtherRegs = 88;
constexpr int kWMaxStages = 5;
constexpr int kNBuf = BSMR_WIDE_NBUF;                       // B stages in flight per producer lane (register ring)
constexpr int kWMaxKChunks = 8;                // K <= 256
constexpr int kWTmemCols = 512;                // kWAccs accumulators of kWCols fp32 columns
constexpr uint32_t kNoCol = 0xFFFFFFFFu;
constexpr int kWEpiRowWords = 36;                              // padded row of the epilogue staging (conflict-free STS.128)
constexpr int kWEpiStageBytes = 32 * kWEpiRowWords * 4;        // 4608 bytes per epilogue warp
constexpr int kWListCap = 384;                                 // work-list entries of a quarter-tile held in shared memory
constexpr int kWListBytes = kWListCap * 8;                     // 3 KB per epilogue warp

struct __align__(16) WideSmemTail {
    uint64_t b_full[kWMaxStages];    // 8 producer warps stored (and fenced) their rows of the stage
    uint64_t b_empty[kWMaxStages];   // the MMAs that read the stage have completed (tcgen05.commit)
    uint64_t a_ready[kWMaxKChunks];  // A image of K-chunk kc stored by the 8 producer warps
    uint64_t a_free;                 // every MMA of the current group has completed: A may be overwritten
    uint64_t tmem_full[kWAccs];
    uint64_t tmem_empty[kWAccs];     // the epilogue warps have read the accumulator
    uint32_t tmem_base;
    uint32_t pad[3];
};

struct WideParams {
    uint32_t K, kchunks, stages;
    uint32_t num_rows;               // reordered (non-empty) rows
    const uint32_t* cta_begin;       // gridDim.x + 1 tile indices: CTA b owns tiles [cta_begin[b], cta_begin[b + 1])
    const uint4* tile_meta;          // {group, first column (offset into cols), #columns, 0}
    const uint32_t* cols;            // distinct columns of the wide groups, ascending inside a group
    const uint32_t* sb_off;          // [(tile * 4 + quarter) * (chunks + 1) + chunk]: first work-list entry of a 32 x 32 sub-block
    const uint2* entries;            // entry: {byte offset inside the staging image (row * 36 + column) * 4, CSR position}
    const uint32_t* reordered_rows;
    const float* A;
    const float* B;
    float* P;
    uint32_t* error_flag;
    unsigned long long* trace;       // optional (tests/perf probes): 32 time stamps per CTA
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define WCLOCK(slot)                                                                          \
    do {                                                                                      \
        if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * 32 + (slot)] = clock64();      \
    } while (0)
#define WTRACE(slot)                                                                          \
    do {                                                                                      \
        if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * 32 + (slot)] = gtime();        \
    } while (0)

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ uint4 rna4(const float4& v) {
    uint4 o;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o.x) : "f"(v.x));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o.y) : "f"(v.y));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o.z) : "f"(v.z));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o.w) : "f"(v.w));
    return o;
}

__global__ void __launch_bounds__(kWThreads, 1)
wide_sddmm_kernel(const WideParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: an integer round trip loses the address space
    // and every access below would become a generic LD/ST instead of LDS/STS (seen in SASS, 3-5x slower)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_img = smem;                                             // kchunks x 16 KB
    uint8_t* b_ring = smem + (size_t)p.kchunks * kWAChunkBytes;        // stages x 32 KB
    uint8_t* epi_stage = b_ring + (size_t)p.stages * kWBStageBytes;    // 4 epilogue warps x 32 rows x 36 words
    uint8_t* epi_lists = epi_stage + (size_t)kWEpiWarps * kWEpiStageBytes;   // 4 x 3 KB
    WideSmemTail* tail = reinterpret_cast<WideSmemTail*>(epi_lists + (size_t)kWEpiWarps * kWListBytes);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t KC = p.kchunks, S = p.stages;
    // tile range of this CTA (host-side partition: CTAs do not straddle row groups when there are enough of them)
    const uint32_t my_begin = __ldg(p.cta_begin + blockIdx.x), my_end = __ldg(p.cta_begin + blockIdx.x + 1);

    if (warp == 0 && lane == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&tail->b_full[s], kWProducerWarps);
            mbar_init(&tail->b_empty[s], 1);
        }
        for (uint32_t k = 0; k < KC; ++k) mbar_init(&tail->a_ready[k], kWProducerWarps);
        mbar_init(&tail->a_free, 1);
        for (int a = 0; a < kWAccs; ++a) {
            mbar_init(&tail->tmem_full[a], 1);
            mbar_init(&tail->tmem_empty[a], kWEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kWMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)), "n"(kWTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;
    if (warp == 0) WTRACE(0);                      // prologue done

    if (warp < kWProducerWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWProducerRegs));
        if (warp == 0) WTRACE(1);                  // registers granted
        // ================= producers =================
        // lane -> (row slot, 16-byte chunk): a warp instruction covers 4 rows x 128 bytes; round r of a stage
        // covers rows r*32 + warp*4 + lane/8.  kNBuf stages of B are in flight per lane (register ring): the
        // bytes in flight per SM (8 warps x 32 lanes x kNBuf x 64 B = 80 KB) are what hides the L2 latency.
        const uint32_t rbase = warp * 4 + (lane >> 3);
        const uint32_t c16 = lane & 7;
        const uint32_t swz = ((c16 ^ (rbase & 7)) << 4);          // (r*32 + rbase) & 7 == rbase & 7
        const uint32_t K = p.K;
        if (my_begin < my_end) {
            float4 buf[kNBuf][kWRounds];
            uint32_t cols_i[kWRounds], cols_n[kWRounds];
            auto load_cols = [&](uint32_t t, uint32_t (&cols)[kWRounds]) {
                const uint4 m = __ldg(p.tile_meta + t);
#pragma unroll
                for (int r = 0; r < kWRounds; ++r) {
                    const uint32_t rr = r * 32 + rbase;
                    cols[r] = rr < m.z ? __ldg(p.cols + m.y + rr) : kNoCol;
                }
            };
            auto issue_b = [&](const uint32_t (&cols)[kWRounds], uint32_t kc, float4 (&v)[kWRounds]) {
#pragma unroll
                for (int r = 0; r < kWRounds; ++r) {
                    v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cols[r] != kNoCol) v[r] = ldg4(p.B + (size_t)cols[r] * K + kc * kWChunk + c16 * 4);
                }
            };
            auto load_arows = [&](uint32_t g, uint32_t (&arow)[4]) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t gi = g * kWRows + r * 32 + rbase;
                    arow[r] = gi < p.num_rows ? __ldg(p.reordered_rows + gi) : kNoCol;
                }
            };
            auto issue_a = [&](const uint32_t (&arow)[4], uint32_t kc, float4* v) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (arow[r] != kNoCol) v[r] = ldg4(p.A + (size_t)arow[r] * K + kc * kWChunk + c16 * 4);
                }
            };
            auto store_a = [&](uint32_t kc, const float4* v) {
                uint8_t* img = a_img + (size_t)kc * kWAChunkBytes;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    *reinterpret_cast<uint4*>(img + (r * 32 + rbase) * 128 + swz) = rna4(v[r]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tail->a_ready[kc]);
            };

            // ---- start-up: column lists, then the whole A image of the first group with every load in flight at once
            // (the B ring registers double as staging: 4 chunks = 16 float4 per pass)
            uint32_t ti = my_begin, kci = 0;             // issue cursor
            uint32_t ts = my_begin, kcs = 0;             // store cursor
            const uint32_t n_items = (my_end - my_begin) * KC;
            uint32_t issued = 0, stored = 0, stage = 0, phase = 0, a_loads = 1;
            uint32_t cur_group = __ldg(p.tile_meta + my_begin).x;
            auto issue_next = [&](float4 (&v)[kWRounds]) {
                if (issued < n_items) {
                    issue_b(cols_i, kci, v);
                    ++issued;
                    if (++kci == KC) {
                        kci = 0;
                        ++ti;
#pragma unroll
                        for (int r = 0; r < kWRounds; ++r) cols_i[r] = cols_n[r];
                        if (ti + 1 < my_end) load_cols(ti + 1, cols_n);
                    }
                }
            };
            {
                uint32_t arow[4];
                load_arows(cur_group, arow);
                load_cols(ti, cols_i);
                if (ti + 1 < my_end) load_cols(ti + 1, cols_n);
                issue_next(buf[0]);                  // the first B stage travels together with the A image
                constexpr int kAPerBuf = kWRounds / 4;              // a K-chunk of A is 4 float4 per lane
                constexpr int kAPass = (kNBuf - 1) * kAPerBuf;      // K-chunks of A staged per pass in buf[1 .. kNBuf-1]
                for (uint32_t k0 = 0; k0 < KC; k0 += kAPass) {
#pragma unroll
                    for (int j = 0; j < kAPass; ++j)
                        if (k0 + j < KC) issue_a(arow, k0 + j, &buf[1 + j / kAPerBuf][(j % kAPerBuf) * 4]);
#pragma unroll
                    for (int j = 0; j < kAPass; ++j)
                        if (k0 + j < KC) store_a(k0 + j, &buf[1 + j / kAPerBuf][(j % kAPerBuf) * 4]);
                }
            }
            if (warp == 0) WTRACE(2);              // A image stored
#pragma unroll
            for (int b = 1; b < kNBuf - 1; ++b) issue_next(buf[b]);
            bool done = false;
            while (!done) {
#pragma unroll
                for (int b = 0; b < kNBuf; ++b) {
                    if (stored == n_items) { done = true; break; }
                    issue_next(buf[(b + kNBuf - 1) % kNBuf]);
                    if (kcs == 0 && ts != my_begin) {
                        const uint32_t g = __ldg(p.tile_meta + ts).x;
                        if (g != cur_group) {
                            // group switch inside a CTA's range (more groups than CTAs): the previous group's MMAs must
                            // have finished reading A before it is overwritten; chunk by chunk (registers are taken)
                            mbar_wait(&tail->a_free, (a_loads - 1) & 1, p.error_flag, 11);
                            uint32_t arow[4];
                            load_arows(g, arow);
                            for (uint32_t kc = 0; kc < KC; ++kc) {
                                float4 v[4];
                                issue_a(arow, kc, v);
                                store_a(kc, v);
                            }
                            cur_group = g;
                            ++a_loads;
                        }
                    }
                    mbar_wait(&tail->b_empty[stage], phase ^ 1, p.error_flag, 12);
                    uint8_t* st = b_ring + (size_t)stage * kWBStageBytes;
#pragma unroll
                    for (int r = 0; r < kWRounds; ++r)
                        *reinterpret_cast<uint4*>(st + (r * 32 + rbase) * 128 + swz) = rna4(buf[b][r]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tail->b_full[stage]);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                    if (warp == 0 && stored < 4) WTRACE(3 + stored);   // first four B stages stored
                    ++stored;
                    if (++kcs == KC) { kcs = 0; ++ts; }
                }
            }
            if (warp == 0) WTRACE(7);              // producer done
        }
    } else if (warp >= kWMmaWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWOtherRegs));
        if (warp == kWMmaWarp) {
        // ================= MMA issuer =================
        uint32_t stage = 0, phase = 0, it = 0, a_idx = 0, cur_group = kNoCol;
        for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
            const uint4 m = __ldg(p.tile_meta + t);
            const uint32_t acc = it % kWAccs, acc_phase = (it / kWAccs) & 1;
            const bool new_group = m.x != cur_group;
            cur_group = m.x;
            const bool last_of_group = (t + 1 == my_end) || (__ldg(p.tile_meta + t + 1).x != m.x);
            const uint32_t n_mma = (m.z + 15u) & ~15u;
            const uint32_t idesc = make_idesc_tf32(kWRows, n_mma);
            mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1, p.error_flag, 13);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * kWCols;
            for (uint32_t kc = 0; kc < KC; ++kc) {
                if (new_group) mbar_wait(&tail->a_ready[kc], a_idx & 1, p.error_flag, 14);
                mbar_wait(&tail->b_full[stage], phase, p.error_flag, 15);
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t da = make_smem_desc(smem_u32(a_img + (size_t)kc * kWAChunkBytes));
                    const uint64_t db = make_smem_desc(smem_u32(b_ring + (size_t)stage * kWBStageBytes));
#pragma unroll
                    for (uint32_t k = 0; k < kWChunk / 8; ++k)
                        umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                    umma_commit(&tail->b_empty[stage]);
                    if (it == 0 && kc == 0) WTRACE(8);             // first MMAs issued
                    if (kc + 1 == KC) {
                        umma_commit(&tail->tmem_full[acc]);
                        if (last_of_group) umma_commit(&tail->a_free);
                    }
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            if (new_group) ++a_idx;
        }
        WTRACE(9);                                 // last MMA issued
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWOtherRegs));
        // ================= epilogue (4 warps) =================
        // Per 32-column chunk: tcgen05.ld of the warp's 32 x 32 accumulator sub-block -> padded shared-memory staging
        // (row stride 36 words: conflict-free 128-bit stores) -> the sub-block's work list, 32 entries per pass: one
        // LDS gather and one STG per entry.  The instruction count follows the nnz, not the 128 x 256 tile area (a
        // predicated store per accumulator element cost 4.6 us per tile, measured).  The quarter-tile's list (position,
        // CSR index) is copied to shared memory with cp.async one tile ahead: read from global memory inside the chunk
        // loop it put one L2 round trip per chunk on the critical path (3.5 us per tile, measured).
        const uint32_t quarter = warp & 3;          // TMEM lanes [32*quarter, +32): fixed by warp id % 4
        float* stg = reinterpret_cast<float*>(epi_stage + (size_t)quarter * kWEpiStageBytes);
        const uint2* lent = reinterpret_cast<const uint2*>(epi_lists + (size_t)quarter * kWListBytes);
        const uint32_t lent_u32 = smem_u32(lent);
        const uint8_t* stg_bytes = reinterpret_cast<const uint8_t*>(stg);
        // lane j < 9 holds the first entry of chunk j of the tile in hand (lane 8: end of chunk 7)
        auto fetch_offsets = [&](uint32_t t) -> uint32_t {
            return lane <= kWWords ? __ldg(p.sb_off + ((size_t)t * 4 + quarter) * (kWWords + 1) + lane) : 0u;
        };
        auto copy_lists = [&](uint32_t offs) {      // entries [E0, E1) of the quarter-tile; E0 is a multiple of 8
            const uint32_t E0 = __shfl_sync(0xffffffffu, offs, 0), E1 = __shfl_sync(0xffffffffu, offs, kWWords);
            uint32_t n = E1 - E0;
            if (n > (uint32_t)kWListCap) n = kWListCap;
            for (uint32_t i = lane * 2; i < n; i += 64)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lent_u32 + i * 8), "l"(p.entries + E0 + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint32_t it = 0;
        uint32_t off_cur = 0, off_next = 0;
        if (my_begin < my_end) {
            off_cur = fetch_offsets(my_begin);
            copy_lists(off_cur);
            if (my_begin + 1 < my_end) off_next = fetch_offsets(my_begin + 1);
        }
        for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
            const uint32_t acc = it % kWAccs, acc_phase = (it / kWAccs) & 1;
            const uint32_t E0 = __shfl_sync(0xffffffffu, off_cur, 0);
            uint32_t eoff[kWWords + 1];             // chunk boundaries relative to the quarter-tile's list
#pragma unroll
            for (int j = 0; j <= kWWords; ++j) eoff[j] = __shfl_sync(0xffffffffu, off_cur, j) - E0;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            mbar_wait(&tail->tmem_full[acc], acc_phase, p.error_flag, 16);
            tc_fence_after();
            if (quarter == 0 && it < 2) WTRACE(10 + 2 * it);   // accumulator of tile 0 / 1 complete
#pragma unroll
            for (int j = 0; j < kWWords; ++j) {
                const uint32_t e0 = eoff[j], e1 = eoff[j + 1];
                if (e0 == e1) continue;
                uint32_t v[32];
                const bool tr = quarter == 0 && it == 1 && j < 3;
                if (tr) WCLOCK(16 + j * 4);
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kWCols + j * 32u;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (tr) WCLOCK(17 + j * 4);
                uint4* srow = reinterpret_cast<uint4*>(stg + lane * kWEpiRowWords);
#pragma unroll
                for (int i = 0; i < 8; ++i) srow[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                __syncwarp();
                if (tr) WCLOCK(18 + j * 4);
                // four passes of 32 entries at a time: the list reads, the gathers and the stores of a group are
                // independent of one another (one epilogue warp per scheduler: every dependent instruction costs its
                // full latency, so the loop is written for instruction count and ILP)
                if (e1 <= (uint32_t)kWListCap) {
                    for (uint32_t eb = e0; eb < e1; eb += 128) {
                        uint2 en[4];
                        float val[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t e = eb + u * 32 + lane;
                            en[u] = lent[e < e1 ? e : e0];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) val[u] = *reinterpret_cast<const float*>(stg_bytes + en[u].x);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (eb + u * 32 + lane < e1) p.P[en[u].y] = val[u];
                    }
                } else {                              // list longer than the shared-memory copy: straight from global memory
                    for (uint32_t e = e0 + lane; e < e1; e += 32) {
                        const uint2 en = e < (uint32_t)kWListCap ? lent[e] : __ldg(p.entries + E0 + e);
                        p.P[en.y] = *reinterpret_cast<const float*>(stg_bytes + en.x);
                    }
                }
                __syncwarp();
                if (tr) WCLOCK(19 + j * 4);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);
            if (quarter == 0 && it < 2) WTRACE(11 + 2 * it);   // epilogue of tile 0 / 1 done
            // next tile: its list goes into the buffer this tile no longer needs
            off_cur = off_next;
            if (t + 1 < my_end) copy_lists(off_cur);
            if (t + 2 < my_end) off_next = fetch_offsets(t + 2);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (quarter == 0) WTRACE(14);              // epilogue done
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) WTRACE(15);
    if (warp == kWMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kWTmemCols));
    }
}

}  // namespace

// CTA -> tile range table for the wide kernel over tiles [tile_begin, tile_end).  When there are at least as many
// CTAs as row groups no CTA straddles a group (an A reload in mid-range costs K/32 serial L2 round trips): every
// group first gets one CTA, the remaining CTAs go one by one to the group with the most tiles per CTA, and a group's
// tiles are split evenly over its CTAs.  Otherwise: equal contiguous ranges.
int wide_partition(bsmr_plan* plan, uint32_t tile_begin, uint32_t tile_end) {
    bsmr_ctx* ctx = plan->ctx;
    plan->w_part_begin = tile_begin;
    plan->w_part_end = tile_end;
    plan->w_grid = 0;
    if (tile_end <= tile_begin) return BSMR_OK;
    const uint32_t ntiles = tile_end - tile_begin;
    const uint32_t ctas = ntiles < (uint32_t)ctx->sm_count ? ntiles : (uint32_t)ctx->sm_count;
    std::vector<uint32_t> gstart;   // first tile of every group in the range (+ end)
    for (uint32_t t = tile_begin; t < tile_end; ++t)
        if (t == tile_begin || plan->h_wt_group[t] != plan->h_wt_group[t - 1]) gstart.push_back(t);
    gstart.push_back(tile_end);
    const uint32_t ngroups = static_cast<uint32_t>(gstart.size()) - 1;
    std::vector<uint32_t> table;
    table.reserve(ctas + 1);
    if (ngroups <= ctas) {
        std::vector<uint32_t> share(ngroups, 1);
        for (uint32_t left = ctas - ngroups; left > 0; --left) {
            uint32_t best = 0;
            double best_load = -1.0;
            for (uint32_t g = 0; g < ngroups; ++g) {
                const double load = static_cast<double>(gstart[g + 1] - gstart[g]) / share[g];
                if (load > best_load) { best_load = load; best = g; }
            }
            if (best_load <= 1.0) break;   // every CTA already has at most one tile
            ++share[best];
        }
        for (uint32_t g = 0; g < ngroups; ++g) {
            const uint32_t n = gstart[g + 1] - gstart[g];
            const uint32_t c = share[g] < n ? share[g] : n;
            for (uint32_t i = 0; i < c; ++i) table.push_back(gstart[g] + static_cast<uint32_t>((static_cast<uint64_t>(n) * i) / c));
        }
    } else {
        for (uint32_t i = 0; i < ctas; ++i) table.push_back(tile_begin + static_cast<uint32_t>((static_cast<uint64_t>(ntiles) * i) / ctas));
    }
    table.push_back(tile_end);
    plan->w_grid = static_cast<uint32_t>(table.size()) - 1;
    BSMR_TRY(plan->w_cta_begin.alloc(table.size()));
    BSMR_CUDA_OK(cudaMemcpyAsync(plan->w_cta_begin.ptr, table.data(), table.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BSMR_OK;
}

static unsigned long long* g_wide_trace = nullptr;
extern "C" void bsmr_debug_set_wide_trace(unsigned long long* device_buffer) { g_wide_trace = device_buffer; }

bool wide_supports(uint32_t K, const float* dA, const float* dB) {
    return K >= 32 && K % kWChunk == 0 && K / kWChunk <= kWMaxKChunks &&
           (reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB)) % 16 == 0;
}

int launch_wide(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t tile_begin, uint32_t tile_end,
                cudaStream_t stream) {
    bsmr_ctx* ctx = plan->ctx;
    if (tile_end <= tile_begin) return BSMR_OK;
    if (!wide_supports(K, dA, dB)) {
        set_error("wide row-group path needs K %% 32 == 0, K <= 256 and 16-byte aligned A/B; K = %u", K);
        return BSMR_ERR_UNSUPPORTED;
    }
    const uint32_t kchunks = K / kWChunk;
    // Shared memory and L1 share 228 KB per SM, and every LDG in flight holds an L1 line: with the full 227 KB of
    // shared memory the producers could keep only ~1 KB of B in flight per SM (measured: 3 us per 32 KB stage).  The
    // ring is therefore short (the register ring is the prefetch pipeline, the smem ring only decouples the stores from
    // the MMAs): 2 stages keep K <= 128 under the 132 KB carve-out (96 KB of L1 for kNBuf x 16 KB in flight).
    const size_t max_smem = 232448;   // 227 KB per CTA on sm_100
    const size_t fixed = 1024 + sizeof(WideSmemTail) + (size_t)kWEpiWarps * (kWEpiStageBytes + kWListBytes) + (size_t)kchunks * kWAChunkBytes;
    uint32_t stages = static_cast<uint32_t>((max_smem - fixed) / kWBStageBytes);
    static const uint32_t stage_cap = [] { const char* e = std::getenv("BSMR_WIDE_STAGES"); return e ? (uint32_t)std::atoi(e) : 0u; }();
    const uint32_t want = stage_cap ? stage_cap : 2u;
    if (stages > want) stages = want;
    if (stages > (uint32_t)kWMaxStages) stages = kWMaxStages;
    const size_t smem = fixed + (size_t)stages * kWBStageBytes;
    static bool attr_set = false;
    if (!attr_set) {
        BSMR_CUDA_OK(cudaFuncSetAttribute(wide_sddmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        attr_set = true;
    }
    static DevBuf<uint32_t> error_flag;
    if (!error_flag.ptr) {
        BSMR_TRY(error_flag.alloc(1));
        BSMR_CUDA_OK(cudaMemsetAsync(error_flag.ptr, 0, 4, ctx->stream));
        BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    WideParams p{};
    p.K = K; p.kchunks = kchunks; p.stages = stages;
    p.num_rows = static_cast<uint32_t>(plan->h_reordered_rows.size());
    p.cta_begin = plan->w_cta_begin.ptr;
    p.tile_meta = plan->wt_meta.ptr;
    p.cols = plan->w_cols.ptr;
    p.sb_off = plan->w_sb_off.ptr;
    p.entries = plan->w_entries.ptr;
    p.reordered_rows = plan->reordered_rows.ptr;
    p.A = dA; p.B = dB; p.P = dP;
    p.error_flag = error_flag.ptr;
    p.trace = g_wide_trace;
    g_wide_trace = nullptr;
    if (plan->w_part_begin != tile_begin || plan->w_part_end != tile_end || plan->w_grid == 0) {
        set_error("launch_wide: no CTA partition for tiles [%u, %u)", tile_begin, tile_end);
        return BSMR_ERR_BAD_STATE;
    }
    const uint32_t grid = plan->w_grid;   // one CTA per SM
    wide_sddmm_kernel<<<grid, kWThreads, smem, stream>>>(p);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

}  // namespace bsmr
