// Dense-block SDDMM kernel for sm_100a: TMA gather -> shared memory -> tcgen05.mma (TF32,
// fp32 accumulate in TMEM) -> tcgen05.ld -> S-mask + scatter to CSR order, all in one kernel.
//
// Replaces sddmm_gpu_dense_block_m16n16k8_matrixA_rowMaj_matrixB_colMaj and its K<=32 twin
// (src/sddmmKernel.cu:213-351, 355-488): wmma m16n16k8, one warp per 16x16 block, operands
// staged with scalar __ldg into padded smem, accumulator fragments scattered through
// blockValues.
//
// Why it is not a transliteration.  tcgen05.mma needs M in {64,128} per CTA; a BSMR row panel
// has only 16 rows and every panel has its own dense column list, so panels cannot be stacked
// along M.  The operands are therefore SWAPPED: one work item ("tile") is up to 128 dense
// columns of one panel (8 reference blocks) and the MMA computes the transposed product
//      D[128 cols x 16 rows] = Bcols[128 x K] * Arows[16 x K]^T
//   * M side = 128 gathered columns of B.  B is column-major (ld = K), so a column is a
//     contiguous K-vector: the matrix [N x K] is "K-major" as tcgen05 wants it.
//   * N side = the panel's 16 gathered rows of A (row-major, K-major as well), N = 16.
//   * K is consumed 32 floats (= 128 bytes = one SWIZZLE_128B atom row) per pipeline stage,
//     4 tcgen05.mma.kind::tf32 (K = 8 each) per stage.
// TMA: cp.async.bulk.tensor.2d ... tile::gather4 fetches 4 arbitrary rows of the [N x K] (or
// [M x K]) tensor per request and lays them down as 4 consecutive 128-byte rows of the
// swizzled smem tile; 32 requests fill the B-column tile, 4 the A-row tile.  Sentinel columns
// (index N, the reference's padding) and rows past the last panel row are out of bounds for
// the tensor map and arrive as zeros.
// Precision: tcgen05 kind::tf32 reads the fp32 bit patterns and IGNORES the low 13 mantissa
// bits (truncation; measured on B200: mean signed error -6.5e-4 on all-positive inputs).  The
// reference rounds to nearest (wmma::__float_to_tf32 = cvt.rna, src/sddmmKernel.cu:317-322).  The rounding is done
// by the TMA unit: the tensor maps carry CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, which makes it round every fp32 element
// to TF32 on the way into shared memory (measured against a cvt.rna converter pass in shared memory, which this
// kernel used before: same maximum error, mean signed error -3e-7; tests/tf32_probe.py).
// Warp roles (13 warps): 0..7 = TMA producers (16 columns of the B-column tile each, see the producer section for
// why so many), 8 = TMEM allocator + MMA issuer, 9..12 = epilogue (tcgen05.ld of the warp's 32 TMEM lanes x 16 columns,
// mask + scatter P[idx] = acc).
// Four TMEM accumulators (4 x 16 columns) let the epilogue of tile i overlap the MMAs of tiles i+1..i+3; a 5-stage
// smem ring (18 KB / stage) keeps ~90 KB of loads in flight per CTA.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace bsmr {
namespace {

using namespace tc;

constexpr int kStages = 5;
constexpr int kChunk = 32;                         // floats of K per stage (128 bytes)
constexpr int kBTileBytes = kTileCols * kChunk * 4;   // 16384
constexpr int kATileBytes = kPanel * kChunk * 4;      // 2048
constexpr int kProducerWarps = 8;                 // each owns 16 dense columns = 4 gather4 requests per stage
constexpr int kColsPerProducer = 128 / kProducerWarps;
constexpr int kReqPerProducer = kColsPerProducer / 4;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kEpiWarp0 = kProducerWarps + 1;     // 4 epilogue warps (kEpiWarp0 % 4 is irrelevant: quarter = warp & 3 covers all four)
constexpr int kDenseThreads = (kProducerWarps + 5) * 32;
constexpr int kAccs = 4;                           // TMEM accumulators in rotation (tile i+4 waits for the epilogue of tile i)
constexpr int kTmemCols = kAccs * 16;              // 16 fp32 columns each

struct __align__(16) DenseSmemTail {
    uint64_t full[kStages];    // TMA bytes landed
    uint64_t empty[kStages];   // MMAs that read the stage have completed
    uint64_t tmem_full[kAccs];
    uint64_t tmem_empty[kAccs];
    uint32_t tmem_base;
    uint32_t pad[3];
};
constexpr size_t kDenseSmemBytes = 1024 /*alignment slack*/ + (size_t)kStages * (kBTileBytes + kATileBytes) + sizeof(DenseSmemTail);

struct DenseParams {
    uint32_t K;
    uint32_t M;              // rows of A (out-of-bounds row index for missing panel rows)
    uint32_t N;
    uint32_t num_rows;       // reordered (non-empty) rows
    uint32_t tile_begin, tile_end;
    const uint32_t* reordered_rows;
    const uint32_t* dense_cols;
    const uint32_t* tile_panel;
    const uint32_t* tile_col_begin;
    const uint32_t* tile_ncols;
    const uint32_t* tile_scatter;
    const uint4* tile_meta;      // {panel, first dense column (offset into dense_cols), #columns, 0} per tile
    const uint32_t* tile_list;   // optional: [tile_begin, tile_end) index this list of tile ids (groups on the wide path removed)
    float* P;
    uint32_t* error_flag;
    // batch (sddmm_gpu_batch, src/sddmmKernel.cu:2764-2848): work item w = batch element * #tiles + tile; the batch's A
    // matrices are one [batch * M, K] tensor, its B matrices one [batch * N, K] tensor (the reference's strides M*K / N*K)
    uint32_t batch;
    uint32_t oob_row, oob_col;   // out-of-bounds coordinates of those tensors: batch * M, batch * N
    size_t stride_p;             // nnz
#ifdef BSMR_DEBUG
    uint32_t* debug_smem;    // optional: raw copy of stage 0 of the first tile (probe / tests)
#endif
};

__global__ void __launch_bounds__(kDenseThreads, 2)
dense_sddmm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const DenseParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: an integer round trip loses the address space
    // and every access below would become a generic LD/ST instead of LDS/STS (seen in SASS, 3-5x slower)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_tiles = smem;                                   // kStages x 16 KB, 1024-aligned
    uint8_t* a_tiles = smem + (size_t)kStages * kBTileBytes;     // kStages x 2 KB, 1024-aligned
    DenseSmemTail* tail = reinterpret_cast<DenseSmemTail*>(a_tiles + (size_t)kStages * kATileBytes);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t num_chunks = (p.K + kChunk - 1) / kChunk;
    const uint32_t ntiles = p.tile_end - p.tile_begin;
    const uint32_t total = ntiles * p.batch;           // work items: (batch element, tile)

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&tail->full[s], kProducerWarps);   // every producer warp arrives (see the producer loop)
            mbar_init(&tail->empty[s], 1);
        }
        for (int a = 0; a < kAccs; ++a) {
            mbar_init(&tail->tmem_full[a], 1);
            mbar_init(&tail->tmem_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp < kProducerWarps) {
        // ================= TMA producers (warps 0..7) =================
        // Producer warp w owns B-tile rows [16w, 16w + 16): lane l < 4 issues the gather4 of dense columns
        // 16w + 4l .. + 3; the last producer warp's lanes 4..7 additionally issue the four gather4 requests of the A-row
        // tile.  Lane 0 of warp 0 arms the stage's mbarrier with the byte count of ALL requests of the stage (a
        // complete_tx that lands before the expect_tx only makes the transaction count transiently negative).
        // Why eight warps: a gather4 takes its coordinates from uniform registers, ptxas serialises the issuing lanes of
        // a warp with an ELECT / R2UR loop at ~140 cycles per request (ncu source view); one warp issuing all 36 requests
        // of a stage needed ~1.7 us per stage, four warps ~0.6 us.  (One lane issuing 8 requests from straight-line code
        // was slower still: measured.)
        uint32_t stage = 0, phase = 0;
        auto fetch = [&](uint32_t w, uint32_t& nc, int4& cols, int4& rows) {
            const uint32_t be = w / ntiles, ti = p.tile_begin + (w - be * ntiles);
            const uint32_t t = p.tile_list ? __ldg(p.tile_list + ti) : ti;
            const uint4 m = __ldg(p.tile_meta + t);
            nc = m.z;
            const int oc = (int)p.oob_col, orow = (int)p.oob_row;
            cols = make_int4(oc, oc, oc, oc);
            const uint32_t c0 = warp * kColsPerProducer + lane * 4;      // first dense column of this lane's request
            if (lane < kReqPerProducer && c0 < nc) {
                const int4 raw = __ldg(reinterpret_cast<const int4*>(p.dense_cols + m.y + c0));
                const int cb = (int)(be * p.N);          // sentinel columns (index N, the reference's padding) stay out of bounds
                cols.x = (uint32_t)raw.x < p.N ? raw.x + cb : oc;
                cols.y = (uint32_t)raw.y < p.N ? raw.y + cb : oc;
                cols.z = (uint32_t)raw.z < p.N ? raw.z + cb : oc;
                cols.w = (uint32_t)raw.w < p.N ? raw.w + cb : oc;
            }
            rows = make_int4(orow, orow, orow, orow);
            if (warp == kProducerWarps - 1 && lane >= 4 && lane < 8) {
                const uint32_t r0 = m.x * kPanel + (lane - 4) * 4;
                int* rp = reinterpret_cast<int*>(&rows);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (r0 + j < p.num_rows) rp[j] = (int)(__ldg(p.reordered_rows + r0 + j) + be * p.M);
            }
        };
        uint32_t nc = 0, nc_next = 0;
        int4 cols, rows, cols_next, rows_next;
        cols = rows = cols_next = rows_next = make_int4(0, 0, 0, 0);
        uint32_t t = blockIdx.x;
        if (t < total) fetch(t, nc, cols, rows);
        for (; t < total; t += gridDim.x) {
            // indices of the next tile are fetched while this one streams (dependent L2/DRAM loads off the critical path)
            if (t + gridDim.x < total) fetch(t + gridDim.x, nc_next, cols_next, rows_next);
            const bool has_cols = lane < kReqPerProducer && warp * kColsPerProducer + lane * 4 < nc;
            const bool has_rows = warp == kProducerWarps - 1 && lane >= 4 && lane < 8;
            const uint32_t tx_bytes = (nc / 4) * 512u + kATileBytes;
            for (uint32_t kc = 0; kc < num_chunks; ++kc) {
                mbar_wait(&tail->empty[stage], phase ^ 1, p.error_flag, 1);
                // every producer warp arrives, with or without requests for this tile: a warp that only watched could fall
                // a whole ring cycle behind, and a parity wait cannot tell phase n from phase n + 2
                if (lane == 0) {
                    if (warp == 0) mbar_arrive_expect_tx(&tail->full[stage], tx_bytes);
                    else mbar_arrive(&tail->full[stage]);
                }
                uint8_t* bt = b_tiles + (size_t)stage * kBTileBytes;
                uint8_t* at = a_tiles + (size_t)stage * kATileBytes;
                const int x = (int)(kc * kChunk);
                if (has_cols) tma_gather4(&map_b, &tail->full[stage], bt + (warp * kReqPerProducer + lane) * 512, x, cols);
                if (has_rows) tma_gather4(&map_a, &tail->full[stage], at + (lane - 4) * 512, x, rows);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            nc = nc_next;
            cols = cols_next;
            rows = rows_next;
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_tf32(kTileCols, kPanel);
        uint32_t stage = 0, phase = 0, it = 0;
        for (uint32_t t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const uint32_t acc = it % kAccs, acc_phase = (it / kAccs) & 1;
            mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1, p.error_flag, 2);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * kPanel;
            for (uint32_t kc = 0; kc < num_chunks; ++kc) {
                mbar_wait<false>(&tail->full[stage], phase, p.error_flag, 3);
                tc_fence_after();
#ifdef BSMR_DEBUG
                if (p.debug_smem && t == 0 && kc == 0 && blockIdx.x == 0) {
                    // probe: raw image of stage 0 (B-column tile then A-row tile)
                    const uint32_t* src_b = reinterpret_cast<const uint32_t*>(b_tiles);
                    const uint32_t* src_a = reinterpret_cast<const uint32_t*>(a_tiles);
                    for (uint32_t i = lane; i < kBTileBytes / 4; i += 32) p.debug_smem[i] = src_b[i];
                    for (uint32_t i = lane; i < kATileBytes / 4; i += 32) p.debug_smem[kBTileBytes / 4 + i] = src_a[i];
                    __syncwarp();
                }
#endif
                if (lane == 0) {
                    const uint64_t da = make_smem_desc(smem_u32(b_tiles + (size_t)stage * kBTileBytes));
                    const uint64_t db = make_smem_desc(smem_u32(a_tiles + (size_t)stage * kATileBytes));
#pragma unroll
                    for (uint32_t k = 0; k < kChunk / 8; ++k) {
                        // advance 8 tf32 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) address field
                        umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&tail->empty[stage]);
                    if (kc + 1 == num_chunks) umma_commit(&tail->tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (four warps after the MMA warp) =================
        const uint32_t quarter = warp & 3;              // TMEM lanes [32*quarter, 32*quarter + 32)
        const uint32_t c = quarter * 32 + lane;         // dense column of the tile owned by this thread
        uint32_t it = 0;
        for (uint32_t w = blockIdx.x; w < total; w += gridDim.x, ++it) {
            const uint32_t be = w / ntiles, ti = p.tile_begin + (w - be * ntiles);
            float* Pb = p.P + be * p.stride_p;
            asm volatile("" : "+l"(Pb));        // one register pair per work item: not re-derived per store
            const uint32_t t = p.tile_list ? __ldg(p.tile_list + ti) : ti;
            const uint32_t acc = it % kAccs, acc_phase = (it / kAccs) & 1;
            const uint32_t nc = __ldg(p.tile_meta + t).z;
            const bool active = quarter * 32 < nc;
            uint32_t idx[kPanel];
            if (active) {      // issued before the accumulator is waited for: the loads overlap the tile's MMAs
                const uint32_t* sc = p.tile_scatter + (size_t)t * kPanel * kTileCols + c;
#pragma unroll
                for (int r = 0; r < (int)kPanel; ++r) idx[r] = __ldg(sc + r * kTileCols);
            }
            mbar_wait(&tail->tmem_full[acc], acc_phase, p.error_flag, 4);
            tc_fence_after();
            uint32_t v[kPanel];
            if (active) {
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kPanel;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);
            if (active) {
#pragma unroll
                for (int r = 0; r < (int)kPanel; ++r)
                    st_global_if(Pb, idx[r] != kNull ? idx[r] : 0u, v[r], idx[r] != kNull);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

}  // namespace

#ifdef BSMR_DEBUG
// probe hook (debug builds only): when set, the next dense launch copies the first tile's stage-0 smem image here
static uint32_t* g_debug_smem = nullptr;
extern "C" void bsmr_debug_set_dense_smem_dump(uint32_t* device_buffer) { g_debug_smem = device_buffer; }
#endif

int launch_dense(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t tile_begin, uint32_t tile_end,
                 const uint32_t* tile_list, cudaStream_t stream, uint32_t batch) {
    bsmr_ctx* ctx = plan->ctx;
    if (tile_end <= tile_begin || batch == 0) return BSMR_OK;
    if (!dense_supports(K, dA, dB)) {
        set_error("dense-block path needs K %% 4 == 0 and 16-byte aligned A/B (TMA row stride); K = %u", K);
        return BSMR_ERR_UNSUPPORTED;
    }
    if ((uint64_t)plan->M * batch > 0x7FFFFFFFull || (uint64_t)plan->N * batch > 0x7FFFFFFFull ||
        (uint64_t)(tile_end - tile_begin) * batch > 0xFFFFFFFFull) {
        set_error("dense-block path: batch %u x (%u rows, %u columns) exceeds the 31-bit TMA coordinates", batch, plan->M, plan->N);
        return BSMR_ERR_UNSUPPORTED;
    }
    CUtensorMap map_a, map_b;
    // TFLOAT32 maps: the TMA unit rounds to TF32.  The batch's matrices are contiguous (stride M*K / N*K): one tensor each.
    BSMR_TRY(make_row_gather_map(ctx, dA, plan->M * batch, K, &map_a, true));
    BSMR_TRY(make_row_gather_map(ctx, dB, plan->N * batch, K, &map_b, true));

    if (!ctx->attr_dense) {       // per context = per device: the attribute is a property of the function ON a device
        BSMR_CUDA_OK(cudaFuncSetAttribute(dense_sddmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDenseSmemBytes));
        ctx->attr_dense = true;
    }
    uint32_t* error_flag = kernel_error_flag();
    if (!error_flag) {
        set_error("no mapped host memory for the kernels' error flag");
        return BSMR_ERR_CUDA;
    }
    DenseParams p{};
    p.K = K; p.M = plan->M; p.N = plan->N;
    p.num_rows = static_cast<uint32_t>(plan->h_reordered_rows.size());
    p.tile_begin = tile_begin; p.tile_end = tile_end;
    p.reordered_rows = plan->reordered_rows.ptr;
    p.dense_cols = plan->dense_cols.ptr;
    p.tile_panel = plan->tile_panel.ptr;
    p.tile_col_begin = plan->tile_col_begin.ptr;
    p.tile_ncols = plan->tile_ncols.ptr;
    p.tile_scatter = plan->tile_scatter.ptr;
    p.tile_meta = plan->tile_meta.ptr;
    p.tile_list = tile_list;
    p.P = dP;
    p.error_flag = error_flag;
    p.batch = batch;
    p.oob_row = plan->M * batch;
    p.oob_col = plan->N * batch;
    p.stride_p = plan->nnz;
#ifdef BSMR_DEBUG
    p.debug_smem = g_debug_smem;
    g_debug_smem = nullptr;
#endif
    const uint64_t items = (uint64_t)(tile_end - tile_begin) * batch;
    const uint32_t max_ctas = static_cast<uint32_t>(ctx->sm_count) * 2;  // 2 CTAs (2 x 94 KB smem, 2 x 32 TMEM columns) per SM
    const uint32_t grid = items < max_ctas ? (uint32_t)items : max_ctas;
    dense_sddmm_kernel<<<grid, kDenseThreads, kDenseSmemBytes, stream>>>(map_a, map_b, p);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

bool dense_supports(uint32_t K, const float* dA, const float* dB) {
    return K % 4 == 0 && (reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB)) % 16 == 0;
}

}  // namespace bsmr
