// Wide row-group SDDMM kernel for sm_100a.  A 256-row group of the reordered matrix stays resident in shared memory as
// two 128-row M operands of tcgen05.mma (TF32, fp32 accumulators in TMEM); the group's distinct B columns stream past
// them 128 at a time as the N operand (one B stage feeds both sub-groups); the epilogue keeps only the accumulator
// elements S has, driven by per-sub-block work lists.
//
// Why it exists (no counterpart in the reference, whose only tensor-core unit is the 16 x 16 block of
// src/sddmmKernel.cu:213-351): on matrices that are dense-ish at the scale of a row group (the nips example is 4 %
// dense, DLMC masks 2-30 %) both reference-shaped kernels are bound by the L2 -> SM gather of one K-vector of B per nnz
// (or per 16-row panel column), which saturates near 5.4 TB/s on B200 (measured with both kernels).  Here a B column
// is fetched once per 256 rows: the gather traffic drops by nnz(group) / distinct_columns(group) (~11x on nips), and
// the work that replaces it -- 256 x 128 x K MACs per tile whatever the fill -- is what the tensor pipe has to spare.
// A group takes this path when nnz(group) >= ratio * (128 * tiles + 256) (colreorder.cu: build_wide_format); every
// other group keeps the BSMR dense-block + residual kernels.
//
// Pipeline (17 warps, one CTA per SM, persistent over a contiguous range of tiles of one row group):
//   warps 0-7   TMA producers: cp.async.bulk.tensor.2d tile::gather4 (4 arbitrary rows of the [N x K] / [M x K] tensor
//               per request, SWIZZLE_128B K-major image, 32 floats of K per stage); first the A images of the group
//               ([sub-group][K-chunk] x 16 KB), then the B ring (S x 16 KB)
//   warps 12-15 converters: cvt.rna.tf32.f32 in place on every landed image (tcgen05 kind::tf32 truncates, the reference
//               rounds to nearest), fence.proxy.async, mbarrier
//   warp  16    TMEM allocator (512 columns) + single-lane MMA issuer: per B stage 4 MMAs (M=128, N<=128, K=8) per
//               resident sub-group; accumulator sets rotate so that the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 8-11  epilogue (TMEM lane quarter = warp % 4): tcgen05.ld 32x32b.x32 -> padded staging -> work list -> P
// Shared memory (K = 128): 128 KB A + 4 x 16 KB B ring + 18 KB staging + 14 KB lists.  A first version staged the
// operands with LDG -> cvt -> STS from producer warps (no converter pass, less shared-memory traffic): every load in
// flight held an L1 line, L1 is what shared memory leaves over, and the epilogue's LDS/STG queued behind the loads in
// the LSU; measured slower at every depth of prefetch.  TMA does not touch the LSU / L1 miss path at all.
// Roofline: HBM on the compulsory bytes of the step; inside, L2 -> SM traffic (distinct columns x K x 4 per group).
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace bsmr {
namespace {

using namespace tc;

constexpr int kWSubRows = 128;                        // UMMA M = TMEM lanes: one sub-group of a row group
constexpr int kWGroupRows = BSMR_WIDE_GROUP_ROWS;     // 256
constexpr int kWSub = kWGroupRows / kWSubRows;        // sub-groups per row group
constexpr int kWCols = BSMR_WIDE_TILE_COLS;           // max columns of a wide tile = UMMA N = 128
constexpr int kWChunk = 32;                           // floats of K per stage (128 bytes = one swizzle row)
constexpr int kWAImgBytes = kWSubRows * 128;          // 16 KB: one sub-group x one K-chunk of A
constexpr int kWBStageBytes = kWCols * 128;           // 16 KB
constexpr int kWWords = kWCols / 32;                  // 32-column quarters of a tile (= TMEM lane quarters)
constexpr int kWRowQ = kWGroupRows / 32;              // 32-row quarters of a row group
constexpr int kWProducerWarps = 8;                    // warps 0-7: TMA gather4 issue
constexpr int kWEpiWarp0 = 8;                         // warps 8-15: epilogue, TMEM lane quarter = warp % 4, row half = (warp - 8) / 4
constexpr int kWEpiWarps = 8;
constexpr int kWMmaWarp = 16;
constexpr int kWThreads = 17 * 32;
constexpr int kWMaxStages = 8;
constexpr int kWMaxKChunks = 8;                       // K <= 256
constexpr int kWTmemCols = 512;
constexpr int kWMaxAccs = 4;
constexpr uint32_t kNoCol = 0xFFFFFFFFu;
constexpr int kWEpiRowWords = 36;                              // padded row of the epilogue staging (conflict-free STS.128)
constexpr int kWEpiStageBytes = 32 * kWEpiRowWords * 4;        // 4608 bytes per epilogue warp
constexpr int kWListPage = 104;                                // work-list entries per page of an epilogue warp's list stream
constexpr int kWListBytes = 2 * kWListPage * 8;                // two pages in shared memory per epilogue warp: 1.6 KB

struct __align__(16) WideSmemTail {
    uint64_t b_full[kWMaxStages];    // TMA bytes of the stage landed
    uint64_t b_empty[kWMaxStages];   // the MMAs that read the stage have completed (tcgen05.commit)
    uint64_t a_full[kWMaxKChunks];   // A images of K-chunk kc landed
    uint64_t a_free;                 // every MMA that reads the current A images has completed
    uint64_t tmem_full[kWMaxAccs];
    uint64_t tmem_empty[kWMaxAccs];  // the epilogue warps have read the accumulator
    uint32_t tmem_base;
    uint32_t pad[3];
};

struct WideParams {
    uint32_t K, kchunks, stages;
    uint32_t sgp;                    // sub-groups per pass: 2 (A images of the whole group resident, K <= 128) or 1 (K = 256)
    uint32_t num_rows;               // reordered (non-empty) rows
    uint32_t M, N;                   // out-of-bounds coordinates for missing rows / columns (TMA zero fill)
    const uint32_t* cta_begin;       // gridDim.x + 1 tile indices: CTA b owns tiles [cta_begin[b], cta_begin[b + 1])
    const uint4* tile_meta;          // {group, first column (offset into cols, multiple of 4), #columns, 0}
    const uint32_t* cols;            // distinct columns of the wide groups, ascending inside a group
    uint32_t num_tiles;              // wide tiles of the plan (stride of the per-quarter list streams)
    const uint32_t* sb_off;          // [(column quarter * num_tiles + tile) * 9 + row quarter]: first work-list entry of a 32 x 32 sub-block
    const uint2* entries;            // entry: {byte offset inside the staging image (column * 36 + row) * 4, CSR position}
    const uint32_t* reordered_rows;
    float* P;
    uint32_t* error_flag;
    uint32_t debug;                  // timing experiments only (BSMR_WIDE_DEBUG): 4 = epilogue skips every chunk (wrong results)
    unsigned long long* trace;       // optional (tests/perf probes): 32 time stamps per CTA
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define WTRACE(slot)                                                                          \
    do {                                                                                      \
        if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * 32 + (slot)] = gtime();        \
    } while (0)

__global__ void __launch_bounds__(kWThreads, 1)
wide_sddmm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const WideParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: an integer round trip loses the address space
    // and every access below would become a generic LD/ST instead of LDS/STS (seen in SASS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t KC = p.kchunks, S = p.stages, SGP = p.sgp;
    uint8_t* a_img = smem;                                                   // [KC][SGP] x 16 KB: per K-chunk one (SGP x 128)-row image
    uint8_t* b_ring = a_img + (size_t)SGP * KC * kWAImgBytes;                // S x 16 KB
    uint8_t* epi_stage = b_ring + (size_t)S * kWBStageBytes;                 // 4 epilogue warps x 32 rows x 36 words
    uint8_t* epi_lists = epi_stage + (size_t)kWEpiWarps * kWEpiStageBytes;   // 4 x 3 KB
    WideSmemTail* tail = reinterpret_cast<WideSmemTail*>(epi_lists + (size_t)kWEpiWarps * kWListBytes);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t passes = kWSub / SGP;                  // K = 256: the tile range is walked once per sub-group
    const uint32_t naccs = kWTmemCols / (kWCols * SGP);   // accumulator sets in rotation (one set = SGP x 128 columns)
    // tile range of this CTA (host-side partition: CTAs do not straddle row groups when there are enough of them)
    const uint32_t my_begin = __ldg(p.cta_begin + blockIdx.x), my_end = __ldg(p.cta_begin + blockIdx.x + 1);

    if (warp == 0 && lane == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&tail->b_full[s], 1);
            mbar_init(&tail->b_empty[s], 1);
        }
        for (uint32_t k = 0; k < KC; ++k) {
            mbar_init(&tail->a_full[k], 1);
        }
        mbar_init(&tail->a_free, 1);
        for (int a = 0; a < kWMaxAccs; ++a) {
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
        // ================= TMA producers (warps 0..7) =================
        // A stage is 32 gather4 requests (4 B columns x 128 bytes each, laid down as 4 consecutive rows of the
        // SWIZZLE_128B K-major image); producer warp w issues requests 4w..4w+3 from its lanes 0..3 (a gather4 takes its
        // coordinates from uniform registers, so ptxas serialises the lanes of a warp: eight warps issue in parallel).
        // Lane 0 of warp 0 arms the stage's mbarrier with the byte count of all its requests.  Missing columns / rows
        // carry the out-of-bounds coordinate and arrive as zeros.  Nothing here touches the LSU / L1 miss path: with
        // LDG-staged operands every load in flight held an L1 line and the epilogue's LDS/STG queued behind them.
        const uint32_t rq = warp * 4 + lane;         // request (= group of 4 rows of the image) of this lane, lanes 0..3
        const bool issuer = lane < 4;
        auto fetch_cols = [&](uint32_t t, uint32_t& ncols, int4& cols) {
            const uint4 m = __ldg(p.tile_meta + t);
            ncols = m.z;
            cols = make_int4((int)p.N, (int)p.N, (int)p.N, (int)p.N);
            const uint32_t c0 = rq * 4;
            if (issuer && c0 < m.z) {
                cols = __ldg(reinterpret_cast<const int4*>(p.cols + m.y + c0));
                if (c0 + 1 >= m.z) cols.y = (int)p.N;
                if (c0 + 2 >= m.z) cols.z = (int)p.N;
                if (c0 + 3 >= m.z) cols.w = (int)p.N;
            }
        };
        uint32_t stage = 0, phase = 0, a_loads = 0, cur_key = kNoCol;
        for (uint32_t pass = 0; pass < passes; ++pass) {
            uint32_t ncols = 0, ncols_next = 0;
            int4 cols = make_int4(0, 0, 0, 0), cols_next = make_int4(0, 0, 0, 0);
            if (my_begin < my_end) fetch_cols(my_begin, ncols, cols);
            for (uint32_t t = my_begin; t < my_end; ++t) {
                if (t + 1 < my_end) fetch_cols(t + 1, ncols_next, cols_next);   // indices of the next tile: off the critical path
                const uint32_t g = __ldg(p.tile_meta + t).x;
                const uint32_t key = g * 2 + pass;
                const bool new_key = key != cur_key;
                int4 arows[2];
                arows[0] = arows[1] = make_int4((int)p.M, (int)p.M, (int)p.M, (int)p.M);
                if (new_key) {
                    // (re)load the A images, [sub-group of the pass][K-chunk] x 32 requests, interleaved with the B stages of
                    // this tile: chunk kc of A, then stage kc of B, so that the first MMAs start after one chunk has landed
                    // instead of after the whole 128 KB image
                    if (a_loads > 0) mbar_wait(&tail->a_free, (a_loads - 1) & 1, p.error_flag, 11);
                    if (issuer) {
                        for (uint32_t sg = 0; sg < SGP; ++sg) {
                            const uint32_t r0 = g * kWGroupRows + (pass * SGP + sg) * kWSubRows + rq * 4;
                            int* rp = reinterpret_cast<int*>(&arows[sg]);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (r0 + j < p.num_rows) rp[j] = (int)__ldg(p.reordered_rows + r0 + j);
                        }
                    }
                    cur_key = key;
                    ++a_loads;
                }
                const bool has_cols = issuer && rq * 4 < ncols;
                const uint32_t tx_bytes = ((ncols + 3) / 4) * 512u;
                for (uint32_t kc = 0; kc < KC; ++kc) {
                    if (new_key) {
                        if (warp == 0 && lane == 0) mbar_arrive_expect_tx(&tail->a_full[kc], SGP * kWAImgBytes);
                        if (issuer) {
                            for (uint32_t sg = 0; sg < SGP; ++sg)
                                tma_gather4(&map_a, &tail->a_full[kc], a_img + ((size_t)kc * SGP + sg) * kWAImgBytes + rq * 512,
                                            (int)(kc * kWChunk), arows[sg]);
                        }
                    }
                    mbar_wait(&tail->b_empty[stage], phase ^ 1, p.error_flag, 12);
                    if (warp == 0 && lane == 0) mbar_arrive_expect_tx(&tail->b_full[stage], tx_bytes);
                    if (has_cols)
                        tma_gather4(&map_b, &tail->b_full[stage], b_ring + (size_t)stage * kWBStageBytes + rq * 512, (int)(kc * kWChunk), cols);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                ncols = ncols_next;
                cols = cols_next;
            }
        }
    } else if (warp == kWMmaWarp) {
        // ================= MMA issuer =================
        uint32_t stage = 0, phase = 0, it = 0, a_idx = 0, cur_key = kNoCol;
        for (uint32_t pass = 0; pass < passes; ++pass) {
            for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
                const uint4 m = __ldg(p.tile_meta + t);
                const uint32_t acc = it % naccs, acc_phase = (it / naccs) & 1;
                const uint32_t key = m.x * 2 + pass;
                const bool new_key = key != cur_key;
                cur_key = key;
                const bool last_of_key = (t + 1 == my_end) || (__ldg(p.tile_meta + t + 1).x != m.x);
                // transposed product: M = the tile's 128 B columns (TMEM lanes), N = the resident rows (SGP x 128 TMEM columns);
                // one MMA reads 4 KB of B and SGP x 4 KB of A for 128 x N x 8 MACs -- 96 B/cycle of shared memory at N = 256
                // instead of the 128 B/cycle (the whole pipe) of two M = 128, N = 128 MMAs, which starved the TMA writes
                const uint32_t idesc = make_idesc_tf32(kWCols, SGP * kWSubRows);
                mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1, p.error_flag, 13);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * (SGP * kWSubRows);
                for (uint32_t kc = 0; kc < KC; ++kc) {
                    if (new_key) mbar_wait<false>(&tail->a_full[kc], a_idx & 1, p.error_flag, 14);
                    mbar_wait<false>(&tail->b_full[stage], phase, p.error_flag, 15);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t da = make_smem_desc(smem_u32(b_ring + (size_t)stage * kWBStageBytes));
                        const uint64_t db = make_smem_desc(smem_u32(a_img + (size_t)kc * SGP * kWAImgBytes));
                        if (!(p.debug & 16u)) {
#pragma unroll
                            for (uint32_t k = 0; k < kWChunk / 8; ++k)
                                umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&tail->b_empty[stage]);
                        if (it == 0 && kc == 0) WTRACE(8);             // first MMAs issued
                        if (kc + 1 == KC) {
                            umma_commit(&tail->tmem_full[acc]);
                            if (last_of_key) umma_commit(&tail->a_free);
                        }
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if (new_key) ++a_idx;
            }
        }
        WTRACE(9);                                 // last MMA issued
    } else {
        // ================= epilogue (warps 8..11) =================
        // Per (tile, sub-group) and 32-column chunk: tcgen05.ld of the warp's 32 x 32 accumulator sub-block -> padded
        // shared-memory staging (row stride 36 words: conflict-free 128-bit stores) -> the sub-block's work list, four
        // passes of 32 entries at a time: one LDS.64 (entry), one LDS (value), one STG per entry.  The instruction count
        // follows the nnz, not the tile area (a predicated store per accumulator element cost 4.6 us per 128 x 256
        // tile, measured: one epilogue warp per scheduler pays every dependent instruction's full latency).  The list of
        // stream of the warp is paged through shared memory with cp.async.
        const uint32_t ew = warp - kWEpiWarp0;
        const uint32_t quarter = warp & 3;          // TMEM lanes [32*quarter, +32) = tile columns: fixed by warp id % 4
        const uint32_t half = ew >> 2;              // rows [128*half, +128) of the group
        float* stg = reinterpret_cast<float*>(epi_stage + (size_t)ew * kWEpiStageBytes);
        const uint8_t* stg_bytes = reinterpret_cast<const uint8_t*>(stg);
        const uint2* lpage = reinterpret_cast<const uint2*>(epi_lists + (size_t)ew * kWListBytes);
        const uint32_t lpage_u32 = smem_u32(lpage);
        // The lists of this warp's units (tile, column quarter, row half), tile ascending, are ONE contiguous stream of
        // entries in global memory.  The warp pages through it: two pages of kWListPage entries in shared memory, page
        // n + 2 requested (cp.async) the moment page n is used up, whatever the fill of the tiles.
        const size_t stream_row = (size_t)(quarter * 2 + half) * p.num_tiles;
        auto unit_offsets = [&](uint32_t t) -> uint32_t {   // lane r <= 4: first entry of row quarter r of the unit (lane 4: end)
            return lane <= 4u ? __ldg(p.sb_off + (stream_row + t) * 5 + lane) : 0u;
        };
        uint32_t stream_base = 0;
        auto request_page = [&](uint32_t pg) {
            const uint2* src = p.entries + stream_base + (size_t)pg * kWListPage;
            const uint32_t dst = lpage_u32 + (pg & 1u) * (kWListPage * 8);
            for (uint32_t i = lane * 2; i < (uint32_t)kWListPage; i += 64)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 8), "l"(src + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint32_t cur_page = 0;
        auto ensure_page = [&](uint32_t pg) {       // pages are consumed in ascending order
            while (cur_page < pg) {
                // page cur_page must have LANDED (a skipped page may still be in flight) and every lane must be done with
                // it before its buffer takes page cur_page + 2
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
                request_page(cur_page + 2);
                ++cur_page;
            }
            asm volatile("cp.async.wait_group 1;" ::: "memory");    // all but the newest request (page cur_page + 1) have landed
            __syncwarp();
        };
        uint32_t it = 0;
        for (uint32_t pass = 0; pass < passes; ++pass) {
            // K = 256: one row half is resident per pass and only the warps of that half have work.  The others still take
            // part in the accumulator hand-shake tile by tile: a warp that skipped ahead would test the parity of a phase
            // the barrier has not reached yet (a parity wait can only tell the current phase from the previous one).
            if (SGP == 1 && half != pass) {
                for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
                    const uint32_t acc = it % naccs, acc_phase = (it / naccs) & 1;
                    mbar_wait(&tail->tmem_full[acc], acc_phase, p.error_flag, 19);
                    if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);
                    __syncwarp();
                }
                continue;
            }
            const uint32_t col0 = SGP == 2 ? half * 128u : 0u;      // first accumulator column of this warp's rows
            uint32_t off_next = my_begin < my_end ? unit_offsets(my_begin) : 0u;
            if (my_begin < my_end) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                stream_base = __shfl_sync(0xffffffffu, off_next, 0);     // multiple of 8 entries: 16-byte aligned pages
                cur_page = 0;
                request_page(0);
                request_page(1);
            }
            for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
                const uint32_t acc = it % naccs, acc_phase = (it / naccs) & 1;
                const uint32_t off_cur = off_next;
                if (t + 1 < my_end) off_next = unit_offsets(t + 1);    // requested now, used one tile later
                uint32_t eoff[5];                   // chunk boundaries relative to the stream
#pragma unroll
                for (int c = 0; c <= 4; ++c) eoff[c] = __shfl_sync(0xffffffffu, off_cur, c) - stream_base;
                mbar_wait(&tail->tmem_full[acc], acc_phase, p.error_flag, 16);
                tc_fence_after();
                if (ew == 0 && it < 2) WTRACE(10 + 2 * it);   // accumulators of tile 0 / 1 complete
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t e0 = eoff[c], e1 = eoff[c + 1];
                    if (e0 == e1 || (p.debug & 4u)) continue;
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * (SGP * kWSubRows) + col0 + c * 32u;
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
                    // staging image [column = lane][row]: thread writes the 32 rows of its column
                    uint4* srow = reinterpret_cast<uint4*>(stg + lane * kWEpiRowWords);
#pragma unroll
                    for (int i = 0; i < 8; ++i) srow[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    __syncwarp();
                    // the chunk's entries, page by page, four groups of 32 entries at a time
                    for (uint32_t e = e0; e < e1;) {
                        const uint32_t pg = e / kWListPage;
                        ensure_page(pg);
                        const uint32_t pend = (pg + 1) * kWListPage;
                        const uint32_t seg_end = e1 < pend ? e1 : pend;
                        const uint2* lent = lpage + (pg & 1u) * kWListPage - (size_t)pg * kWListPage;   // lent[e] = entry e of the stream
                        for (uint32_t eb = e; eb < seg_end; eb += 128) {
                            uint2 en[4];
                            float val[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t ee = eb + q * 32 + lane;
                                en[q] = lent[ee < seg_end ? ee : e];
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) val[q] = *reinterpret_cast<const float*>(stg_bytes + en[q].x);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (eb + q * 32 + lane < seg_end) p.P[en[q].y] = val[q];
                        }
                        e = seg_end;
                    }
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);
                if (ew == 0 && it < 2) WTRACE(11 + 2 * it);   // epilogue of tile 0 / 1 done
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (ew == 0) WTRACE(14);                   // epilogue done
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
    // the A images of both sub-groups stay resident when they fit (K <= 128: 2 x K/32 x 16 KB <= 128 KB); at K = 256
    // one sub-group is resident at a time and the CTA walks its tile range twice
    const uint32_t sgp = kchunks <= 4 ? 2u : 1u;
    const size_t max_smem = 232448;   // 227 KB per CTA on sm_100
    const size_t fixed = 1024 + sizeof(WideSmemTail) + (size_t)kWEpiWarps * (kWEpiStageBytes + kWListBytes) +
                         (size_t)sgp * kchunks * kWAImgBytes;
    uint32_t stages = static_cast<uint32_t>((max_smem - fixed) / kWBStageBytes);
    static const uint32_t stage_cap = [] { const char* e = std::getenv("BSMR_WIDE_STAGES"); return e ? (uint32_t)std::atoi(e) : 0u; }();
    if (stage_cap && stages > stage_cap) stages = stage_cap;
    if (stages > (uint32_t)kWMaxStages) stages = kWMaxStages;
    const size_t smem = fixed + (size_t)stages * kWBStageBytes;
    static bool attr_set = false;
    if (!attr_set) {
        BSMR_CUDA_OK(cudaFuncSetAttribute(wide_sddmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        attr_set = true;
    }
    uint32_t* error_flag = kernel_error_flag();
    if (!error_flag) {
        set_error("no mapped host memory for the kernels' error flag");
        return BSMR_ERR_CUDA;
    }
    CUtensorMap map_a, map_b;
    static const bool fp32_maps = std::getenv("BSMR_WIDE_FP32_MAPS") != nullptr;   // experiment: truncating operands
    BSMR_TRY(make_row_gather_map(ctx, dA, plan->M, K, &map_a, !fp32_maps));
    BSMR_TRY(make_row_gather_map(ctx, dB, plan->N, K, &map_b, !fp32_maps));
    WideParams p{};
    p.K = K; p.kchunks = kchunks; p.stages = stages; p.sgp = sgp;
    p.num_rows = static_cast<uint32_t>(plan->h_reordered_rows.size());
    p.M = plan->M; p.N = plan->N;
    p.cta_begin = plan->w_cta_begin.ptr;
    p.tile_meta = plan->wt_meta.ptr;
    p.cols = plan->w_cols.ptr;
    p.num_tiles = plan->num_wide_tiles;
    p.sb_off = plan->w_sb_off.ptr;
    p.entries = plan->w_entries.ptr;
    p.reordered_rows = plan->reordered_rows.ptr;
    p.P = dP;
    p.error_flag = error_flag;
    static const uint32_t dbg = [] { const char* e = std::getenv("BSMR_WIDE_DEBUG"); return e ? (uint32_t)std::atoi(e) : 0u; }();
    p.debug = dbg;
    p.trace = g_wide_trace;
    g_wide_trace = nullptr;
    if (plan->w_part_begin != tile_begin || plan->w_part_end != tile_end || plan->w_grid == 0) {
        set_error("launch_wide: no CTA partition for tiles [%u, %u)", tile_begin, tile_end);
        return BSMR_ERR_BAD_STATE;
    }
    const uint32_t grid = plan->w_grid;   // one CTA per SM
    wide_sddmm_kernel<<<grid, kWThreads, smem, stream>>>(map_a, map_b, p);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

}  // namespace bsmr
