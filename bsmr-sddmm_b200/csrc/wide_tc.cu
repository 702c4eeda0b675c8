// Wide row-group SDDMM kernel for sm_100a.  A 256-row group of the reordered matrix stays resident in shared memory as
// the N operand of tcgen05.mma (TF32, fp32 accumulators in TMEM); the group's distinct B columns stream past it 128 at a
// time as the M operand (transposed product: TMEM lanes = tile columns, TMEM columns = group rows); the epilogue keeps
// only the accumulator elements S has.
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
//   warps 0-7   TMA producers: the A images of the group ([K-chunk][sub-group] x 16 KB, tile::gather4 of the reordered
//               rows), then the B ring (S x 16 KB, 32 floats of K per stage): one tiled load per run of consecutive
//               columns (whole tile or 32-column quarter), tile::gather4 (4 arbitrary rows per request) otherwise.  The
//               tensor maps are CU_TENSOR_MAP_DATA_TYPE_TFLOAT32: the TMA unit rounds fp32 -> TF32 to nearest on the way in
//               (tcgen05 kind::tf32 alone truncates; the reference rounds: src/sddmmKernel.cu:317-322)
//   warp  16    TMEM allocator (512 columns) + single-lane MMA issuer: per B stage 4 MMAs (M = 128 tile columns,
//               N = 256 resident rows, K = 8); two accumulator sets rotate so that the epilogue of tile i overlaps the
//               MMAs of tile i+1.  K = 256: one 128-row half resident at a time, the tile range is walked twice
//   warps 8-15  epilogue (TMEM lane quarter = warp % 4, row half = (warp - 8) / 4), in one of two forms chosen per plan by
//               the fill of the tiles (template parameter): per-entry lists through a staged accumulator image, or row
//               masks with direct stores; both stream their metadata with cp.async.bulk rings (see the epilogue section)
// Shared memory (K = 128): 128 KB A + 16 KB B stages (2 with lists: 36 KB staging + 27 KB list pages; 4 with masks:
// 32 KB row-meta pages).  A first version staged the operands with LDG -> cvt -> STS from producer warps: every load in
// flight held an L1 line, L1 is what shared memory leaves over, and the epilogue's LDS/STG queued behind the loads in
// the LSU; measured slower at every depth of prefetch.  TMA does not touch the LSU / L1 miss path at all.
// Roofline: HBM on the compulsory bytes of the step; inside, L2 -> SM traffic (distinct columns x K x 4 per group) and
// the shared-memory / issue cost of the epilogue.
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
// ---- epilogue, list form (sparse tiles)
constexpr int kWSbRows = kWideSbRows;                          // rows of an epilogue sub-block (32 columns x 32 rows)
static_assert(kWSbRows == 32, "the epilogue loads a sub-block with tcgen05.ld 32x32b.x32");
constexpr int kWUSB = (kWGroupRows / 2) / kWSbRows;            // sub-blocks per unit (tile x column quarter x row half): 4
constexpr int kWEpiRowWords = kWideStagePitchWords;            // pitch of the staging image [32 columns][32 rows] (conflict-free STS.128)
constexpr int kWEpiStageBytes = 32 * kWEpiRowWords * 4;        // 4608 bytes per epilogue warp
constexpr int kWListPage = 216;                                // 8-byte slots per page of an epilogue warp's list stream
constexpr int kWListPages = 2;                                 // pages in shared memory per epilogue warp (bulk-copy ring)
constexpr int kWListBytes = kWListPages * kWListPage * 8;      // 1.7 KB per epilogue warp
constexpr int kWHeaderSlots = 4;                               // a unit's list starts with 8 words: entries in sub-blocks 0..s
// ---- epilogue, mask form (dense tiles)
constexpr int kWUnitRows = kWGroupRows / 2;                    // rows of an epilogue unit (tile x column quarter x row half)
constexpr int kWMetaPageBytes = kWUnitRows * 8;                // one unit = 128 row-meta pairs {mask, first CSR position} = 1 KB
constexpr int kWMetaPages = 4;                                 // units in shared memory per epilogue warp (bulk-copy ring)
constexpr int kWMetaBytes = kWMetaPages * kWMetaPageBytes;     // 4 KB per epilogue warp

static_assert(kWListPages <= kWMetaPages, "l_full is sized for the larger ring");

struct __align__(16) WideSmemTail {
    uint64_t b_full[kWMaxStages];    // TMA bytes of the stage landed
    uint64_t b_empty[kWMaxStages];   // the MMAs that read the stage have completed (tcgen05.commit)
    uint64_t a_full[kWMaxKChunks];   // A images of K-chunk kc landed
    uint64_t a_free;                 // every MMA that reads the current A images has completed
    uint64_t tmem_full[kWMaxAccs];
    uint64_t tmem_empty[kWMaxAccs];  // the epilogue warps have read the accumulator
    uint64_t l_full[kWEpiWarps][kWMetaPages];   // a page of an epilogue warp's list / row-meta stream has landed (bulk copy)
    uint32_t tmem_base;
    uint32_t pad[3];
};

struct WideParams {
    uint32_t K, kchunks, stages;
    uint32_t sgp;                    // sub-groups per pass: 2 (A images of the whole group resident, K <= 128) or 1 (K = 256)
    uint32_t num_rows;               // reordered (non-empty) rows
    uint32_t M, N;                   // rows of one A / columns of one B (row / column offset of a batch element)
    uint32_t oob_row, oob_col;       // out-of-bounds coordinates for missing rows / columns (TMA zero fill): batch * M, batch * N
    uint32_t batch;                  // batch elements: the CTA walks its tile range once per (element, pass); the batch's A / B
                                     // matrices are one [batch * M, K] / [batch * N, K] tensor (sddmm_gpu_batch's strides)
    size_t stride_p;                 // nnz: P of batch element b starts at P + b * stride_p
    const uint4* cta_rec;            // per CTA: {first tile, end tile, group of the first tile, its first column id}, {its tile_meta}
    const uint4* tile_meta;          // {group, first column (offset into cols, multiple of 4), #columns, 0}
    const uint32_t* cols;            // distinct columns of the wide groups, ascending inside a group
    uint32_t num_tiles;              // wide tiles of the plan (stride of the per-quarter list streams)
    const uint2* entries;            // mask form: row-meta pairs [(column quarter * 2 + row half) * num_tiles + tile][128]: {mask, first CSR
                                     // position}; list form: the warps' list streams of 8-byte slots (colreorder.cu)
    const uint32_t* sb_off;          // list form: [(column quarter * 2 + row half) * num_tiles + tile] (+1): first slot of the unit
    const uint32_t* reordered_rows;
    float* P;
    uint32_t* error_flag;
#ifdef BSMR_DEBUG
    unsigned long long* trace;       // probe builds only (tests/wide_trace.py): 32 time stamps per CTA
#endif
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifdef BSMR_DEBUG
#define WTRACE(slot)                                                                          \
    do {                                                                                      \
        if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * 32 + (slot)] = gtime();        \
    } while (0)
#else
#define WTRACE(slot) do { } while (0)
#endif

template <bool kMaskEpilogue>
__global__ void __launch_bounds__(kWThreads, 1)
wide_sddmm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_b32, const __grid_constant__ CUtensorMap map_b128, const WideParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: an integer round trip loses the address space
    // and every access below would become a generic LD/ST instead of LDS/STS (seen in SASS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t KC = p.kchunks, S = p.stages, SGP = p.sgp;
    uint8_t* a_img = smem;                                                   // [KC][SGP] x 16 KB: per K-chunk one (SGP x 128)-row image
    uint8_t* b_ring = a_img + (size_t)SGP * KC * kWAImgBytes;                // S x 16 KB
    // epilogue scratch: list form = staging images + list pages, mask form = row-meta pages
    uint8_t* epi_stage = b_ring + (size_t)S * kWBStageBytes;
    uint8_t* epi_lists = kMaskEpilogue ? epi_stage : epi_stage + (size_t)kWEpiWarps * kWEpiStageBytes;
    WideSmemTail* tail = reinterpret_cast<WideSmemTail*>(epi_lists + (size_t)kWEpiWarps * (kMaskEpilogue ? kWMetaBytes : kWListBytes));

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t passes = kWSub / SGP;                  // K = 256: the tile range is walked once per sub-group
    const uint32_t rounds = p.batch * passes;             // ... and once per batch element: round r = (element r / passes, pass r % passes)
    const uint32_t naccs = kWTmemCols / (kWCols * SGP);   // accumulator sets in rotation (one set = SGP x 128 columns)
    // tile range of this CTA (host-side partition: CTAs do not straddle row groups when there are enough of them)
    const uint4 rec0 = __ldg(p.cta_rec + 2 * blockIdx.x), rec1 = __ldg(p.cta_rec + 2 * blockIdx.x + 1);
    const uint32_t my_begin = rec0.x, my_end = rec0.y;

    if (warp == 0 && lane == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&tail->b_full[s], kWProducerWarps);   // every producer warp arrives (see the producer loop)
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
        for (int w = 0; w < kWEpiWarps; ++w)
            for (int k = 0; k < kWMetaPages; ++k) mbar_init(&tail->l_full[w][k], 1);
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
        // Runs of consecutive columns (flags from the format builder, tile_meta.w) are fetched with ONE tiled load per
        // 32-column quarter, or per tile, instead of 8 / 32 gather4 requests: every request passes through the MIO queue
        // of its warp's scheduler, and with 32 of them per stage the epilogue warps' LDS / STS waited ~700-1000 cycles
        // behind them (measured with clock64 stamps).
        uint32_t col_off = 0, row_off = 0;           // column / row offset of the current batch element in the stacked tensors
        auto fetch_cols = [&](uint32_t t, uint32_t& ncols, uint32_t& flags, int4& cols) {
            const uint4 m = t == my_begin ? rec1 : __ldg(p.tile_meta + t);      // the first tile's meta came with the start record
            ncols = m.z;
            flags = m.w;
            const int oc = (int)p.oob_col;
            cols = make_int4(oc, oc, oc, oc);
            const uint32_t c0 = rq * 4;
            if (issuer && c0 < m.z) {
                cols = __ldg(reinterpret_cast<const int4*>(p.cols + m.y + c0));
                cols.x += (int)col_off;
                cols.y = c0 + 1 >= m.z ? oc : cols.y + (int)col_off;
                cols.z = c0 + 2 >= m.z ? oc : cols.z + (int)col_off;
                cols.w = c0 + 3 >= m.z ? oc : cols.w + (int)col_off;
            }
        };
        const uint32_t myq = warp >> 1;              // 32-column quarter of the tile this warp's requests belong to
        uint32_t stage = 0, phase = 0, a_loads = 0, cur_key = kNoCol;
        for (uint32_t round = 0; round < rounds; ++round) {
            const uint32_t be = round / passes, pass = round - be * passes;
            col_off = be * p.N;
            row_off = be * p.M;
            uint32_t ncols = 0, ncols_next = 0, flags = 0, flags_next = 0;
            int4 cols = make_int4(0, 0, 0, 0), cols_next = make_int4(0, 0, 0, 0);
            if (my_begin < my_end) fetch_cols(my_begin, ncols, flags, cols);
            for (uint32_t t = my_begin; t < my_end; ++t) {
                if (t + 1 < my_end) fetch_cols(t + 1, ncols_next, flags_next, cols_next);   // indices of the next tile: off the critical path
                const uint32_t g = t == my_begin ? rec0.z : __ldg(p.tile_meta + t).x;
                const uint32_t key = g * rounds + round;          // one A image per (row group, batch element, pass)
                const bool new_key = key != cur_key;
                // (two named vectors, not an array indexed by the sub-group: an indexed array ends up in local memory)
                int4 arows0 = make_int4((int)p.oob_row, (int)p.oob_row, (int)p.oob_row, (int)p.oob_row), arows1 = arows0;
                if (new_key) {
                    // (re)load the A images, [sub-group of the pass][K-chunk] x 32 requests, interleaved with the B stages of
                    // this tile: chunk kc of A, then stage kc of B, so that the first MMAs start after one chunk has landed
                    // instead of after the whole 128 KB image
                    if (a_loads > 0) mbar_wait(&tail->a_free, (a_loads - 1) & 1, p.error_flag, 11);
                    if (issuer) {
                        const uint32_t r0 = g * kWGroupRows + pass * SGP * kWSubRows + rq * 4;
                        if (r0 + 0 < p.num_rows) arows0.x = (int)(__ldg(p.reordered_rows + r0 + 0) + row_off);
                        if (r0 + 1 < p.num_rows) arows0.y = (int)(__ldg(p.reordered_rows + r0 + 1) + row_off);
                        if (r0 + 2 < p.num_rows) arows0.z = (int)(__ldg(p.reordered_rows + r0 + 2) + row_off);
                        if (r0 + 3 < p.num_rows) arows0.w = (int)(__ldg(p.reordered_rows + r0 + 3) + row_off);
                        if (SGP == 2) {
                            const uint32_t r1 = r0 + kWSubRows;
                            if (r1 + 0 < p.num_rows) arows1.x = (int)(__ldg(p.reordered_rows + r1 + 0) + row_off);
                            if (r1 + 1 < p.num_rows) arows1.y = (int)(__ldg(p.reordered_rows + r1 + 1) + row_off);
                            if (r1 + 2 < p.num_rows) arows1.z = (int)(__ldg(p.reordered_rows + r1 + 2) + row_off);
                            if (r1 + 3 < p.num_rows) arows1.w = (int)(__ldg(p.reordered_rows + r1 + 3) + row_off);
                        }
                    }
                    cur_key = key;
                    ++a_loads;
                }
                // what this lane issues per stage, and the bytes the stage's barrier has to see
                const bool whole = (flags & 16u) != 0;
                const bool qrun = ((flags >> myq) & 1u) != 0;
                const bool do_whole = whole && warp == 0 && lane == 0;
                const bool do_qrun = !whole && qrun && (warp & 1u) == 0 && lane == 0;
                const bool do_gather = !whole && !qrun && issuer && rq * 4 < ncols;
                uint32_t tx_bytes = 0;
                if (whole) {
                    tx_bytes = kWBStageBytes;
                } else {
#pragma unroll
                    for (uint32_t q = 0; q < (uint32_t)kWWords; ++q) {
                        const uint32_t nq = ncols > q * 32 ? (ncols - q * 32 < 32u ? ncols - q * 32 : 32u) : 0u;
                        tx_bytes += ((flags >> q) & 1u) ? 4096u : ((nq + 3) / 4) * 512u;
                    }
                }
                for (uint32_t kc = 0; kc < KC; ++kc) {
                    if (new_key) {
                        if (warp == 0 && lane == 0) mbar_arrive_expect_tx(&tail->a_full[kc], SGP * kWAImgBytes);
                        if (issuer) {
                            tma_gather4(&map_a, &tail->a_full[kc], a_img + (size_t)kc * SGP * kWAImgBytes + rq * 512, (int)(kc * kWChunk), arows0);
                            if (SGP == 2)
                                tma_gather4(&map_a, &tail->a_full[kc], a_img + ((size_t)kc * SGP + 1) * kWAImgBytes + rq * 512, (int)(kc * kWChunk), arows1);
                        }
                    }
                    mbar_wait(&tail->b_empty[stage], phase ^ 1, p.error_flag, 12);
                    // every producer warp arrives on the stage's barrier, whether it has requests to issue for this tile or
                    // not: a warp that only watched could otherwise fall a whole ring cycle behind the others, and a parity
                    // wait cannot tell phase n from phase n + 2 (seen as a rare hang once a single lane issued a whole stage)
                    if (lane == 0) {
                        if (warp == 0) mbar_arrive_expect_tx(&tail->b_full[stage], tx_bytes);
                        else mbar_arrive(&tail->b_full[stage]);
                    }
                    uint8_t* bst = b_ring + (size_t)stage * kWBStageBytes;
                    if (do_whole) {
                        // the first tile's first column id came with the start record: no wait for the column list
                        if (t == my_begin) tma_load_2d(&map_b128, &tail->b_full[stage], bst, (int)(kc * kWChunk), (int)(rec0.w + col_off));
                        else tma_load_2d(&map_b128, &tail->b_full[stage], bst, (int)(kc * kWChunk), cols.x);
                    }
                    if (do_qrun) tma_load_2d(&map_b32, &tail->b_full[stage], bst + myq * 4096, (int)(kc * kWChunk), cols.x);
                    if (do_gather) tma_gather4(&map_b, &tail->b_full[stage], bst + rq * 512, (int)(kc * kWChunk), cols);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                ncols = ncols_next;
                flags = flags_next;
                cols = cols_next;
            }
        }
    } else if (warp == kWMmaWarp) {
        // ================= MMA issuer =================
        uint32_t stage = 0, phase = 0, it = 0, a_idx = 0, cur_key = kNoCol;
        for (uint32_t round = 0; round < rounds; ++round) {
            for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
                const uint4 m = __ldg(p.tile_meta + t);
                const uint32_t acc = it % naccs, acc_phase = (it / naccs) & 1;
                const uint32_t key = m.x * rounds + round;
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
#pragma unroll
                        for (uint32_t k = 0; k < kWChunk / 8; ++k)
                            umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
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
        // ================= epilogue (warps 8..15) =================
        const uint32_t ew = warp - kWEpiWarp0;
        if constexpr (kMaskEpilogue) {
            // ---- mask form ----
            // Warp = (TMEM lane quarter = 32 tile columns, row half of the group).  tcgen05.ld 32x32b.x32 gives lane c the
            // accumulators of column c for 32 rows; per row the warp reads ONE pair {mask of the row's nnz among its 32
            // columns, CSR position of the first} (a uniform-address LDS.64: a broadcast) and lane c stores its value to
            // P[base + popc(mask & lanes below c)] when bit c is set: a row's entries inside a column quarter are consecutive
            // CSR positions, so the stores of a row coalesce.  No staging of the accumulator tile through shared memory and
            // no per-entry list: earlier versions moved the whole 128 KB tile through STS/LDS and walked {offset, position}
            // lists (LDS.64 -> LDS -> STG per entry), which cost ~2.5 us per tile against ~1 us of MMAs -- every step of
            // that chain has ~100 cycles of latency under load (clock64 stamps, tests/wide_trace.py).
            // The pairs of a (tile, quarter, half) unit are 1 KB; the units a warp walks are contiguous in global memory and
            // are pulled into a ring of kWMetaPages units with cp.async.bulk, kWMetaPages - 1 tiles ahead.
            const uint32_t quarter = warp & 3;          // TMEM lanes [32*quarter, +32) = tile columns: fixed by warp id % 4
            const uint32_t half = ew >> 2;              // rows [128*half, +128) of the group
            const uint8_t* lring = epi_lists + (size_t)ew * kWMetaBytes;
            const uint32_t lring_u32 = smem_u32(lring);
            uint64_t* lfull = tail->l_full[ew];
            const uint2* units = p.entries + (size_t)(quarter * 2 + half) * p.num_tiles * kWUnitRows;   // [tile][128]
            const uint32_t lane_bit = 1u << lane, lanes_below = lane_bit - 1u;
            uint32_t gp = 0;                            // units requested so far (ring position)
            uint32_t it = 0;
            for (uint32_t round = 0; round < rounds; ++round) {
                const uint32_t be = round / passes, pass = round - be * passes;
                float* Pout = p.P + be * p.stride_p;
                asm volatile("" : "+l"(Pout));      // one register pair for the whole round: not re-derived from (element, stride) per store
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
                const uint32_t gp0 = gp;                                 // ring position of tile my_begin
                auto issue_unit = [&](uint32_t t) {      // tile t -> ring buffer (gp0 + t - my_begin) % kWMetaPages
                    if (lane == 0) {
                        const uint32_t g = gp0 + (t - my_begin);
                        uint64_t* bar = &lfull[g % kWMetaPages];
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer's last readers were generic-proxy loads
                        mbar_arrive_expect_tx(bar, kWMetaPageBytes);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(lring_u32 + (g % kWMetaPages) * kWMetaPageBytes), "l"(units + (size_t)t * kWUnitRows),
                                       "r"((uint32_t)kWMetaPageBytes), "r"(smem_u32(bar))
                                     : "memory");
                    }
                };
                for (uint32_t t = my_begin; t < my_end && t < my_begin + (uint32_t)kWMetaPages; ++t) issue_unit(t);
                for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
                    const uint32_t acc = it % naccs, acc_phase = (it / naccs) & 1;
                    const uint32_t g = gp0 + (t - my_begin);
                    const uint8_t* page = lring + (g % kWMetaPages) * kWMetaPageBytes;
                    mbar_wait(&lfull[g % kWMetaPages], (g / kWMetaPages) & 1, p.error_flag, 20);
                    mbar_wait(&tail->tmem_full[acc], acc_phase, p.error_flag, 16);
                    tc_fence_after();
                    if (ew == 0 && it < 2) WTRACE(10 + 2 * it);   // accumulators of tile 0 / 1 complete
                    // four sub-blocks of 32 rows; not unrolled (the body is ~250 instructions)
    #pragma unroll 1
                    for (uint32_t c = 0; c < 4; ++c) {
                        const uint2* rows = reinterpret_cast<const uint2*>(page) + c * 32;
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
                        // rows in two batches of 16 pairs (uniform address: one broadcast wavefront per two rows)
                        uint2 mb[16];
#pragma unroll
                        for (int r = 0; r < 16; ++r) mb[r] = rows[r];
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int r = 0; r < 16; ++r)
                            st_global_if(Pout, mb[r].y + __popc(mb[r].x & lanes_below), v[r], (mb[r].x & lane_bit) != 0);
#pragma unroll
                        for (int r = 0; r < 16; ++r) mb[r] = rows[16 + r];
#pragma unroll
                        for (int r = 0; r < 16; ++r)
                            st_global_if(Pout, mb[r].y + __popc(mb[r].x & lanes_below), v[16 + r], (mb[r].x & lane_bit) != 0);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);
                    if (ew == 0 && it < 2) WTRACE(11 + 2 * it);   // epilogue of tile 0 / 1 done
                    if (t + kWMetaPages < my_end) issue_unit(t + kWMetaPages);   // every lane is past its reads of this buffer
                }
                gp = gp0 + (my_end - my_begin);
            }
        } else {
            // ---- list form ----
            // Warp = (TMEM lane quarter = 32 tile columns, row half of the group).  Per tile it walks 8 sub-blocks of 32 columns
            // x 16 rows: tcgen05.ld 32x32b.x16 -> padded shared-memory staging image [column][row] (pitch 20 words:
            // conflict-free STS.128) -> the sub-block's work list: one LDS.64 (entry), one LDS (value), one STG per entry, so
            // the instruction count follows the nnz, not the tile area (a predicated store per accumulator element cost
            // 4.6 us per 128 x 256 tile, measured).
            // The lists of this warp's units, tile ascending, are ONE contiguous stream of 8-byte slots in global memory
            // (colreorder.cu); the warp pages through it with cp.async.bulk into a ring of kWListPages pages, each with its
            // own mbarrier.  (Paging with cp.async / LDG went through the LSU and whatever L1 the 227 KB of shared memory
            // leave: every list access then waited ~700 cycles behind those loads -- measured with clock64 stamps.)
            const uint32_t quarter = warp & 3;          // TMEM lanes [32*quarter, +32) = tile columns: fixed by warp id % 4
            const uint32_t half = ew >> 2;              // rows [128*half, +128) of the group
            float* stg = reinterpret_cast<float*>(epi_stage + (size_t)ew * kWEpiStageBytes);
            const uint8_t* stg_bytes = reinterpret_cast<const uint8_t*>(stg);
            const uint8_t* lring = epi_lists + (size_t)ew * kWListBytes;
            const uint32_t lring_u32 = smem_u32(lring);
            uint64_t* lfull = tail->l_full[ew];
            const uint32_t* ustart = p.sb_off + (size_t)(quarter * 2 + half) * p.num_tiles;
            uint32_t gp_base = 0;                       // pages issued in earlier passes (ring position of local page 0)
            uint32_t it = 0;
            for (uint32_t round = 0; round < rounds; ++round) {
                const uint32_t be = round / passes, pass = round - be * passes;
                float* Pout = p.P + be * p.stride_p;
                asm volatile("" : "+l"(Pout));      // one register pair for the whole round: not re-derived from (element, stride) per store
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
                if (my_begin >= my_end) continue;
                const uint32_t col0 = SGP == 2 ? half * 128u : 0u;      // first accumulator column of this warp's rows
                const uint32_t s_begin = __ldg(ustart + my_begin), s_end = __ldg(ustart + my_end);   // slots of this warp's stream
                const uint32_t npages = (s_end - s_begin + kWListPage - 1) / kWListPage;
                auto issue_page = [&](uint32_t pg) {     // local page pg -> ring buffer (gp_base + pg) % kWListPages
                    if (lane == 0) {
                        const uint32_t g = gp_base + pg;
                        const uint32_t first = pg * kWListPage;
                        const uint32_t n = s_end - s_begin - first < (uint32_t)kWListPage ? s_end - s_begin - first : (uint32_t)kWListPage;
                        uint64_t* bar = &lfull[g % kWListPages];
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer's last readers were generic-proxy loads
                        mbar_arrive_expect_tx(bar, n * 8u);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(lring_u32 + (g % kWListPages) * (kWListPage * 8)), "l"(p.entries + s_begin + first), "r"(n * 8u),
                                       "r"(smem_u32(bar))
                                     : "memory");
                    }
                };
                auto wait_page = [&](uint32_t pg) {
                    const uint32_t g = gp_base + pg;
                    mbar_wait(&lfull[g % kWListPages], (g / kWListPages) & 1, p.error_flag, 20);
                };
                for (uint32_t pg = 0; pg < npages && pg < (uint32_t)kWListPages; ++pg) issue_page(pg);
                uint32_t cur_page = 0;
                bool cur_landed = false;                 // cur_page has been waited for (an mbarrier poll costs ~200 cycles: only once per page)
                auto ensure_page = [&](uint32_t pg) {    // pages are consumed in ascending order
                    if (pg == cur_page && cur_landed) return;
                    while (cur_page < pg) {
                        // leaving cur_page: it has landed (it may never have been read), every lane is done with it, and its
                        // buffer takes page cur_page + kWListPages
                        if (!cur_landed) wait_page(cur_page);
                        __syncwarp();
                        if (cur_page + kWListPages < npages) issue_page(cur_page + kWListPages);
                        ++cur_page;
                        cur_landed = false;
                    }
                    wait_page(pg);
                    cur_landed = true;
                };
                // slot e of the stream (relative to s_begin) inside the ring
                auto slot_addr = [&](uint32_t e) -> const uint8_t* {
                    const uint32_t pg = e / kWListPage;
                    return lring + ((gp_base + pg) % kWListPages) * (kWListPage * 8) + (e - pg * kWListPage) * 8;
                };
                uint32_t pos = 0;                        // first slot of the current unit
                for (uint32_t t = my_begin; t < my_end; ++t, ++it) {
                    const uint32_t acc = it % naccs, acc_phase = (it / naccs) & 1;
                    // the unit's header: the cumulative sub-block counts (4 slots, never straddles a page: pages and units start
                    // at multiples of 8 slots)
                    ensure_page(pos / kWListPage);
                    const uint32_t hw = lane < 8u ? reinterpret_cast<const uint32_t*>(slot_addr(pos))[lane] : 0u;
                    const uint32_t ebase = pos + kWHeaderSlots;
                    const uint32_t total = __shfl_sync(0xffffffffu, hw, kWUSB - 1);
                    mbar_wait(&tail->tmem_full[acc], acc_phase, p.error_flag, 16);
                    tc_fence_after();
                    if (ew == 0 && it < 2) WTRACE(10 + 2 * it);   // accumulators of tile 0 / 1 complete
                    // not unrolled: eight copies of this body are ~50 KB of SASS, and eight warps walking different copies kept
                    // missing the instruction cache (the body took >1000 cycles per sub-block, measured)
                    uint32_t e1 = ebase;
    #pragma unroll 1
                    for (uint32_t c = 0; c < (uint32_t)kWUSB; ++c) {
                        const uint32_t e0 = e1;
                        e1 = ebase + __shfl_sync(0xffffffffu, hw, c);
                        if (e0 == e1) continue;
                        // the first (up to) 64 entries of the sub-block are requested BEFORE the accumulator is staged: every step
                        // of this chain (LDTM, STS, LDS.64, LDS, STG) has ~100 cycles of latency under load, and the entry fetch
                        // does not depend on the staging image
                        const uint32_t pg0 = e0 / kWListPage;
                        ensure_page(pg0);
                        const uint32_t pend0 = (pg0 + 1) * kWListPage;
                        uint32_t seg0 = e1 < pend0 ? e1 : pend0;
                        if (seg0 > e0 + 128u) seg0 = e0 + 128u;
                        uint2 en0[4];
                        {
                            const uint2* lent = reinterpret_cast<const uint2*>(slot_addr(pg0 * kWListPage)) - (size_t)pg0 * kWListPage;   // lent[e] = slot e
    #pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t ee = e0 + q * 32 + lane;
                                en0[q] = lent[ee < seg0 ? ee : e0];
                            }
                        }
                        uint32_t v[32];
                        const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * (SGP * kWSubRows) + col0 + c * kWSbRows;
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
                        {
                            float val[4];
    #pragma unroll
                            for (int q = 0; q < 4; ++q)
                                val[q] = *reinterpret_cast<const float*>(stg_bytes + en0[q].x);
    #pragma unroll
                            for (int q = 0; q < 4; ++q)
                                st_global_if(Pout, en0[q].y, __float_as_uint(val[q]), e0 + q * 32 + lane < seg0);
                        }
                        // the rest of a long list (more than 128 entries, or a page boundary inside the first 128)
                        for (uint32_t e = seg0; e < e1;) {
                            const uint32_t pg = e / kWListPage;
                            ensure_page(pg);
                            const uint32_t pend = (pg + 1) * kWListPage;
                            const uint32_t seg_end = e1 < pend ? e1 : pend;
                            const uint2* lent = reinterpret_cast<const uint2*>(slot_addr(pg * kWListPage)) - (size_t)pg * kWListPage;
                            for (uint32_t eb = e; eb < seg_end; eb += 64) {
                                uint2 en[2];
                                float val[2];
    #pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const uint32_t ee = eb + q * 32 + lane;
                                    en[q] = lent[ee < seg_end ? ee : e];
                                }
    #pragma unroll
                                for (int q = 0; q < 2; ++q)
                                    val[q] = *reinterpret_cast<const float*>(stg_bytes + en[q].x);
    #pragma unroll
                                for (int q = 0; q < 2; ++q)
                                    st_global_if(Pout, en[q].y, __float_as_uint(val[q]), eb + q * 32 + lane < seg_end);
                            }
                            e = seg_end;
                        }
                        __syncwarp();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);
                    if (ew == 0 && it < 2) WTRACE(11 + 2 * it);   // epilogue of tile 0 / 1 done
                    pos += (kWHeaderSlots + total + 7u) & ~7u;
                }
                // every page that was requested must have landed before the ring is reused or the CTA exits
                for (uint32_t pg = cur_page; pg < npages && pg < cur_page + (uint32_t)kWListPages; ++pg) wait_page(pg);
                gp_base += npages;
            }
        }
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

namespace {
// start records: one dependent load instead of three (CTA range -> tile -> column ids) before a CTA's first TMA request;
// with a cold L2 every link of that chain is a DRAM round trip
__global__ void wide_cta_rec_kernel(uint32_t ctas, const uint32_t* __restrict__ cta_begin, const uint4* __restrict__ meta,
                                    const uint32_t* __restrict__ cols, uint4* __restrict__ rec) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= ctas) return;
    const uint32_t t0 = cta_begin[b], t1 = cta_begin[b + 1];
    const uint4 m = t0 < t1 ? meta[t0] : make_uint4(0, 0, 0, 0);
    rec[2 * b] = make_uint4(t0, t1, m.x, (t0 < t1 && m.z) ? cols[m.y] : 0u);
    rec[2 * b + 1] = m;
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
    BSMR_TRY(plan->w_cta_rec.alloc(static_cast<size_t>(plan->w_grid) * 2));
    wide_cta_rec_kernel<<<(plan->w_grid + 127) / 128, 128, 0, ctx->stream>>>(plan->w_grid, plan->w_cta_begin.ptr, plan->wt_meta.ptr, plan->w_cols.ptr,
                                                                           plan->w_cta_rec.ptr);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BSMR_OK;
}

#ifdef BSMR_DEBUG
static unsigned long long* g_wide_trace = nullptr;
extern "C" void bsmr_debug_set_wide_trace(unsigned long long* device_buffer) { g_wide_trace = device_buffer; }
#endif

bool wide_supports(uint32_t K, const float* dA, const float* dB) {
    return K >= 32 && K % kWChunk == 0 && K / kWChunk <= kWMaxKChunks &&
           (reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB)) % 16 == 0;
}

int launch_wide(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t tile_begin, uint32_t tile_end,
                cudaStream_t stream, uint32_t batch) {
    bsmr_ctx* ctx = plan->ctx;
    if (tile_end <= tile_begin || batch == 0) return BSMR_OK;
    if ((uint64_t)plan->M * batch > 0x7FFFFFFFull || (uint64_t)plan->N * batch > 0x7FFFFFFFull) {
        set_error("wide row-group path: batch %u x (%u rows, %u columns) exceeds the 31-bit TMA coordinates", batch, plan->M, plan->N);
        return BSMR_ERR_UNSUPPORTED;
    }
    if (!wide_supports(K, dA, dB)) {
        set_error("wide row-group path needs K %% 32 == 0, K <= 256 and 16-byte aligned A/B; K = %u", K);
        return BSMR_ERR_UNSUPPORTED;
    }
    const uint32_t kchunks = K / kWChunk;
    // the A images of both sub-groups stay resident when they fit (K <= 128: 2 x K/32 x 16 KB <= 128 KB); at K = 256
    // one sub-group is resident at a time and the CTA walks its tile range twice
    const uint32_t sgp = kchunks <= 4 ? 2u : 1u;
    const size_t max_smem = 232448;   // 227 KB per CTA on sm_100
    const size_t fixed = 1024 + sizeof(WideSmemTail) + (size_t)kWEpiWarps * (plan->wide_mask_epilogue ? (size_t)kWMetaBytes : (size_t)(kWEpiStageBytes + kWListBytes)) +
                         (size_t)sgp * kchunks * kWAImgBytes;
    uint32_t stages = static_cast<uint32_t>((max_smem - fixed) / kWBStageBytes);
    if (stages > (uint32_t)kWMaxStages) stages = kWMaxStages;
    const size_t smem = fixed + (size_t)stages * kWBStageBytes;
    if (!ctx->attr_wide) {        // per context = per device: the attribute is a property of the function ON a device
        BSMR_CUDA_OK(cudaFuncSetAttribute(wide_sddmm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        BSMR_CUDA_OK(cudaFuncSetAttribute(wide_sddmm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        ctx->attr_wide = true;
    }
    uint32_t* error_flag = kernel_error_flag();
    if (!error_flag) {
        set_error("no mapped host memory for the kernels' error flag");
        return BSMR_ERR_CUDA;
    }
    // TFLOAT32 maps (the TMA unit rounds to nearest); the batch's matrices are contiguous (stride M*K / N*K): one tensor each
    CUtensorMap map_a, map_b;
    BSMR_TRY(make_row_gather_map(ctx, dA, (uint64_t)plan->M * batch, K, &map_a, true));
    BSMR_TRY(make_row_gather_map(ctx, dB, (uint64_t)plan->N * batch, K, &map_b, true));
    CUtensorMap map_b32, map_b128;    // tiled loads of 32 / 128 consecutive columns of B
    BSMR_TRY(make_row_gather_map(ctx, dB, (uint64_t)plan->N * batch, K, &map_b32, true, 32));
    BSMR_TRY(make_row_gather_map(ctx, dB, (uint64_t)plan->N * batch, K, &map_b128, true, 128));
    WideParams p{};
    p.K = K; p.kchunks = kchunks; p.stages = stages; p.sgp = sgp;
    p.num_rows = static_cast<uint32_t>(plan->h_reordered_rows.size());
    p.M = plan->M; p.N = plan->N;
    p.oob_row = plan->M * batch; p.oob_col = plan->N * batch;
    p.batch = batch;
    p.stride_p = plan->nnz;
    p.cta_rec = plan->w_cta_rec.ptr;
    p.tile_meta = plan->wt_meta.ptr;
    p.cols = plan->w_cols.ptr;
    p.num_tiles = plan->num_wide_tiles;
    p.entries = plan->w_entries.ptr;
    p.sb_off = plan->w_sb_off.ptr;
    p.reordered_rows = plan->reordered_rows.ptr;
    p.P = dP;
    p.error_flag = error_flag;
#ifdef BSMR_DEBUG
    p.trace = g_wide_trace;
    g_wide_trace = nullptr;
#endif
    if (plan->w_part_begin != tile_begin || plan->w_part_end != tile_end || plan->w_grid == 0) {
        set_error("launch_wide: no CTA partition for tiles [%u, %u)", tile_begin, tile_end);
        return BSMR_ERR_BAD_STATE;
    }
    const uint32_t grid = plan->w_grid;   // one CTA per SM
    if (plan->wide_mask_epilogue) wide_sddmm_kernel<true><<<grid, kWThreads, smem, stream>>>(map_a, map_b, map_b32, map_b128, p);
    else wide_sddmm_kernel<false><<<grid, kWThreads, smem, stream>>>(map_a, map_b, map_b32, map_b128, p);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

}  // namespace bsmr
