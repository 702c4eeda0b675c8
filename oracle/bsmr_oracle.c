/* TEST INFRASTRUCTURE ONLY -- see bsmr_oracle.h.  Plain C restatement of the reference
 * algorithm; each function cites the reference file:line it follows.  Nothing in the product
 * path (bsmr-sddmm_b200/csrc, libbsmr_b200.so) links or calls this file.
 */
#define _GNU_SOURCE
#include "bsmr_oracle.h"

#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ===================================================================================== */
/* a1: .mtx loader.  Follows src/Matrix.cpp:399-480:                                     */
/*   - skip leading lines that start with '%' (:412)                                     */
/*   - header "rows cols nnz" (:414)                                                     */
/*   - each following non-empty line "row col [value]", 1-based, missing value -> 0      */
/*     (getOneLineThreeData, :374-397; words split on ' ', '\t', '\r', util.hpp:176-190)  */
/*   - reject: more than nnz entries (:433), fewer (:446), out-of-range (:452),           */
/*     duplicate coordinate (:457), nnz <= 1 (:462)                                       */
/*   - stable sort by row only (:467-470) -> in-row order is file order                   */
/*   - row offsets by getCsrRowOffsets (:236-250)                                         */
/* ===================================================================================== */
static const char* next_word(const char* p, char* buf, size_t cap) {
    size_t n = 0;
    while (*p && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') {
        if (n + 1 < cap) buf[n++] = *p;
        ++p;
    }
    buf[n] = 0;
    while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
    return p;
}

typedef struct { uint32_t r, c; } rc_pair;
static int cmp_rc(const void* a, const void* b) {
    const rc_pair* x = (const rc_pair*)a; const rc_pair* y = (const rc_pair*)b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return 0;
}

int oracle_load_mtx(const char* path, oracle_csr* out) {
    memset(out, 0, sizeof(*out));
    FILE* f = fopen(path, "r");
    if (!f) return 0;
    char* line = NULL; size_t cap = 0; ssize_t len;
    char w[128];
    int have_header = 0;
    while ((len = getline(&line, &cap, f)) >= 0) {
        if (line[0] == '%') continue;
        have_header = 1;
        break;
    }
    if (!have_header) { free(line); fclose(f); return 0; }
    {
        const char* p = line;
        p = next_word(p, w, sizeof w); out->rows = (uint32_t)atoi(w);
        p = next_word(p, w, sizeof w); out->cols = (uint32_t)atoi(w);
        p = next_word(p, w, sizeof w); out->nnz = (uint32_t)strtod(w, NULL);
    }
    const uint32_t nnz = out->nnz;
    uint32_t* ri = (uint32_t*)malloc(sizeof(uint32_t) * (nnz ? nnz : 1));
    uint32_t* ci = (uint32_t*)malloc(sizeof(uint32_t) * (nnz ? nnz : 1));
    float* va = (float*)malloc(sizeof(float) * (nnz ? nnz : 1));
    uint32_t idx = 0; int ok = 1;
    while ((len = getline(&line, &cap, f)) >= 0) {
        if (len == 0 || line[0] == '\n' || line[0] == 0) continue; /* empty line: skipped (:375) */
        const char* p = line;
        p = next_word(p, w, sizeof w); const uint32_t r = (uint32_t)atoi(w);
        p = next_word(p, w, sizeof w); const uint32_t c = (uint32_t)atoi(w);
        p = next_word(p, w, sizeof w);
        float v = 0.0f;
        if (w[0]) { /* std::stod throws out_of_range on ERANGE and the reference then stores 0 (:383-388) */
            errno = 0;
            const double d = strtod(w, NULL);
            v = errno == ERANGE ? 0.0f : (float)d;
        }
        if (idx >= nnz) { ok = 0; break; }
        ri[idx] = r - 1; ci[idx] = c - 1; va[idx] = v; ++idx;
    }
    free(line); fclose(f);
    if (ok && idx < nnz) ok = 0;
    if (ok) {
        for (uint32_t i = 0; i < nnz; ++i)
            if (ri[i] >= out->rows || ci[i] >= out->cols) { ok = 0; break; }
    }
    if (ok) { /* duplicate check (std::set in the reference; sort + adjacent compare here) */
        rc_pair* pr = (rc_pair*)malloc(sizeof(rc_pair) * (nnz ? nnz : 1));
        for (uint32_t i = 0; i < nnz; ++i) { pr[i].r = ri[i]; pr[i].c = ci[i]; }
        qsort(pr, nnz, sizeof(rc_pair), cmp_rc);
        for (uint32_t i = 1; i < nnz; ++i)
            if (pr[i].r == pr[i - 1].r && pr[i].c == pr[i - 1].c) { ok = 0; break; }
        free(pr);
    }
    if (ok && nnz <= 1) ok = 0;
    if (!ok) { free(ri); free(ci); free(va); memset(out, 0, sizeof(*out)); return 0; }

    /* stable sort by row = counting sort that keeps file order inside a row */
    out->row_offsets = (uint32_t*)calloc((size_t)out->rows + 1, sizeof(uint32_t));
    out->col_indices = (uint32_t*)malloc(sizeof(uint32_t) * nnz);
    out->values = (float*)malloc(sizeof(float) * nnz);
    for (uint32_t i = 0; i < nnz; ++i) out->row_offsets[ri[i] + 1]++;
    for (uint32_t r = 0; r < out->rows; ++r) out->row_offsets[r + 1] += out->row_offsets[r];
    uint32_t* cursor = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)out->rows + 1));
    memcpy(cursor, out->row_offsets, sizeof(uint32_t) * ((size_t)out->rows + 1));
    for (uint32_t i = 0; i < nnz; ++i) {
        const uint32_t dst = cursor[ri[i]]++;
        out->col_indices[dst] = ci[i];
        out->values[dst] = va[i];
    }
    free(cursor); free(ri); free(ci); free(va);
    return 1;
}

void oracle_free_csr(oracle_csr* csr) {
    free(csr->row_offsets); free(csr->col_indices); free(csr->values);
    memset(csr, 0, sizeof(*csr));
}

/* ===================================================================================== */
/* a2: makeData (src/Matrix.cpp:117-138): default-seeded std::mt19937 (seed 5489) fed to   */
/* std::uniform_real_distribution<float>(0, 2).  libstdc++'s generate_canonical<float,24>  */
/* draws ONE 32-bit word per value: float(word) / 2^32, clamped below 1, then * 2 + 0.     */
/* Single-thread stream only (the reference's OpenMP loop races on the shared engine).     */
/* ===================================================================================== */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;
static void mt_seed(mt19937_t* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}
static uint32_t mt_next(mt19937_t* s) {
    if (s->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
void oracle_make_data(size_t n, float* out) {
    mt19937_t s; mt_seed(&s, 5489u);
    for (size_t i = 0; i < n; ++i) {
        float canon = (float)mt_next(&s) / 4294967296.0f;
        if (canon >= 1.0f) canon = nextafterf(1.0f, 0.0f);
        out[i] = canon * 2.0f + 0.0f;
    }
}

/* ===================================================================================== */
/* a14: sddmm_cpu (src/host.cpp:44-76): P[idx] = sum_k A[row,k] * B[k,col], fp32,          */
/* sequential accumulate from 0, k ascending, NO multiplication by S's value.             */
/* A row-major (ld = K), B col-major (ld = K)  (getOneValueForMultiplication,             */
/* src/Matrix.cpp:198-222).                                                               */
/* ===================================================================================== */
void oracle_sddmm_cpu(uint32_t M, uint32_t N, uint32_t K, const float* A, const float* B,
                      const uint32_t* row_offsets, const uint32_t* col_indices, int num_threads,
                      float* P) {
    (void)N;
#ifdef _OPENMP
    const int saved = omp_get_max_threads();
    if (num_threads > 0) omp_set_num_threads(num_threads);
#else
    (void)num_threads;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < (int64_t)M; ++row) {
        const float* a = A + (size_t)row * K;
        for (uint32_t idx = row_offsets[row]; idx < row_offsets[row + 1]; ++idx) {
            const float* b = B + (size_t)col_indices[idx] * K;
            float val = 0.0f;
            for (uint32_t k = 0; k < K; ++k) val += a[k] * b[k];
            P[idx] = val;
        }
    }
#ifdef _OPENMP
    omp_set_num_threads(saved);
#endif
}

/* include/checkData.hpp:21-30 */
int oracle_check_one(float a, float b) {
    const float abs_eps = 1e-5f;
    const float thr = (float)1e-3; /* const float ERROR_THRESHOLD_EPSILON = 1e-3 */
    const float d = fabsf(a - b);
    if (d < abs_eps) return 1;
    float mx = fabsf(a) > fabsf(b) ? fabsf(a) : fabsf(b);
    if (thr > mx) mx = thr;
    return (d / mx) < thr;
}
uint64_t oracle_check_data(uint64_t n, const float* a, const float* b) {
    uint64_t e = 0;
    for (uint64_t i = 0; i < n; ++i) e += !oracle_check_one(a[i], b[i]);
    return e;
}

/* ===================================================================================== */
/* a3: calculateBlockSize (src/rowReordering.cu:1009-1025)                                 */
/* ===================================================================================== */
uint32_t oracle_calculate_block_size(uint32_t M, uint32_t N, uint64_t free_mem_bytes) {
    const uint32_t max_smem = 49152u; /* include/TensorCoreConfig.cuh:14 */
    /* (size_t)M*M*sizeof(UIN) is an integer product, the divisor a float (static_cast<float>(freeMem/2)) */
    const float gm = (float)((uint64_t)M * M * 4u) / (float)(free_mem_bytes / 2);
    const float sm = (float)((uint64_t)N * 4u) / (float)(max_smem / 2);
    const uint32_t a = (uint32_t)ceilf(gm);
    const uint32_t b = (uint32_t)ceilf(sm);
    const uint32_t bs = a > b ? a : b;
    return bs > 16 ? bs : 16;
}

/* ===================================================================================== */
/* a4: calculateDispersion (src/rowReordering.cu:49-93)                                    */
/*   encoding[b]  = #nnz of the row in column block b (block = col / block_size)          */
/*   dispersion   = sum_{b: enc_b>0} (block_size - enc_b)  +  nnz_row * #{b: enc_b>0}      */
/*   empty rows keep encoding 0 / dispersion 0 (kernel returns early, buffers memset 0)   */
/* all u32 arithmetic (wrapping); CTA = 128 threads = 4 warps -> reduction is complete.    */
/* ===================================================================================== */
void oracle_dispersion(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                       uint32_t block_size, uint32_t* encodings, uint32_t* dispersions) {
    const uint32_t nb = (uint32_t)ceilf((float)N / (float)block_size); /* :1035 */
    memset(encodings, 0, sizeof(uint32_t) * (size_t)M * nb);
    for (uint32_t r = 0; r < M; ++r) {
        uint32_t* enc = encodings + (size_t)r * nb;
        const uint32_t nz = row_offsets[r + 1] - row_offsets[r];
        dispersions[r] = 0;
        if (nz == 0) continue;
        for (uint32_t i = row_offsets[r]; i < row_offsets[r + 1]; ++i) enc[col_indices[i] / block_size]++;
        uint32_t res = 0, dense = 0;
        for (uint32_t b = 0; b < nb; ++b) {
            if (enc[b]) { dense++; res += block_size - enc[b]; }
        }
        dispersions[r] = res + nz * dense;
    }
}

/* src/rowReordering.cu:911-920 */
uint32_t oracle_clustering_blockdim(uint32_t nb) {
    if (nb < 32) return 32;
    /* WARP_SIZE * ceil((float)(nb / 4) / 32): nb/4 is an INTEGER division in the reference */
    int cand = 32 * (int)ceilf((float)(nb / 4) / 32.0f);
    if (cand < 32) cand = 32;
    return cand > 1024 ? 1024u : (uint32_t)cand;
}

/* ===================================================================================== */
/* Block reduction as the device performs it (include/cudaUtil.cuh:13-45):                 */
/*   1. xor-shuffle butterfly inside each warp (w = 1,2,4,8,16; ret += partner)            */
/*   2. lane 0 of each warp stores to shm[warp]                                            */
/*   3. for (stride = blockDim/64; stride >= 1; stride >>= 1) warps < stride add           */
/*      shm[warp + stride]   -- for a non-power-of-two warp count this DROPS warps         */
/*      (e.g. 7 warps: result = (w0+w3)+(w1+w4)).                                          */
/* exact != 0: pad the warp count to a power of two so every warp is summed.               */
/* ===================================================================================== */
static uint32_t block_reduce_u32(const uint32_t* part, uint32_t bd, int exact) {
    const uint32_t nw = bd / 32;
    uint32_t shm[64];
    memset(shm, 0, sizeof shm);
    for (uint32_t w = 0; w < nw; ++w) {
        uint32_t s = 0;
        for (uint32_t l = 0; l < 32; ++l) s += part[w * 32 + l];
        shm[w] = s;
    }
    uint32_t stride = nw / 2;
    if (exact) { uint32_t p = 1; while (p < nw) p <<= 1; stride = p / 2; }
    for (; stride >= 1; stride >>= 1)
        for (uint32_t w = 0; w < stride; ++w) shm[w] += shm[w + stride];
    return shm[0];
}

static float block_reduce_f32(const float* part, uint32_t bd, int exact) {
    const uint32_t nw = bd / 32;
    float shm[64];
    for (int i = 0; i < 64; ++i) shm[i] = 0.0f;
    for (uint32_t w = 0; w < nw; ++w) {
        float v[32], t[32];
        memcpy(v, part + w * 32, sizeof v);
        for (uint32_t x = 1; x < 32; x <<= 1) {
            for (uint32_t l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ x];
            memcpy(v, t, sizeof v);
        }
        shm[w] = v[0];
    }
    uint32_t stride = nw / 2;
    if (exact) { uint32_t p = 1; while (p < nw) p <<= 1; stride = p / 2; }
    for (; stride >= 1; stride >>= 1)
        for (uint32_t w = 0; w < stride; ++w) shm[w] = shm[w] + shm[w + stride];
    return shm[0];
}

/* calculate_similarity_norm_weighted_jaccard, UIN overload (src/rowReordering.cu:235-293):  */
/*   thread t owns blocks i = t, t+bd, t+2bd, ... (strided partial sums, ascending i)        */
/*   squares: int e = enc[i]; sum(u32) += e*e   (wrapping)                                   */
/*   both sums 0 -> 1.0 ; exactly one 0 -> 0.0                                               */
/*   norm = sqrtf((float)sum) ; term = (float)enc / norm (IEEE div)                          */
/*   min_sum += fminf(..), max_sum += fmaxf(..) ; sim = min_sum / max_sum                    */
float oracle_similarity(const uint32_t* enc_rep, const uint32_t* enc_cmp, uint32_t nb, uint32_t bd,
                        int exact) {
    uint32_t sq_rep[1024], sq_cmp[1024];
    float mn[1024], mx[1024];
    memset(sq_rep, 0, sizeof(uint32_t) * bd);
    memset(sq_cmp, 0, sizeof(uint32_t) * bd);
    for (uint32_t i = 0; i < nb; ++i) {
        const uint32_t t = i % bd;
        sq_rep[t] += enc_rep[i] * enc_rep[i];
        sq_cmp[t] += enc_cmp[i] * enc_cmp[i];
    }
    const uint32_t s_rep = block_reduce_u32(sq_rep, bd, exact);
    const uint32_t s_cmp = block_reduce_u32(sq_cmp, bd, exact);
    if (s_rep == 0 && s_cmp == 0) return 1.0f;
    if (s_rep == 0 || s_cmp == 0) return 0.0f;
    const float n_rep = sqrtf((float)s_rep);
    const float n_cmp = sqrtf((float)s_cmp);
    for (uint32_t t = 0; t < bd; ++t) { mn[t] = 0.0f; mx[t] = 0.0f; }
    for (uint32_t i = 0; i < nb; ++i) {
        const uint32_t t = i % bd;
        const float a = (float)enc_rep[i] / n_rep;
        const float b = (float)enc_cmp[i] / n_cmp;
        mn[t] = mn[t] + fminf(a, b);
        mx[t] = mx[t] + fmaxf(a, b);
    }
    const float min_sum = block_reduce_f32(mn, bd, exact);
    const float max_sum = block_reduce_f32(mx, bd, exact);
    return min_sum / max_sum;
}

static int cmp_u32_fwd(const void* a, const void* b) {
    const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : (x > y);
}

/* stable ascending argsort on u32 keys (thrust host sort_by_key == stable) */
typedef struct { uint32_t key, pos; } kp_t;
static int cmp_kp(const void* a, const void* b) {
    const kp_t* x = (const kp_t*)a; const kp_t* y = (const kp_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}
static void stable_argsort_u32(const uint32_t* keys, uint32_t n, uint32_t* order, uint32_t* sorted_keys) {
    kp_t* kp = (kp_t*)malloc(sizeof(kp_t) * (n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) { kp[i].key = keys[i]; kp[i].pos = i; }
    qsort(kp, n, sizeof(kp_t), cmp_kp);
    for (uint32_t i = 0; i < n; ++i) { order[i] = kp[i].pos; if (sorted_keys) sorted_keys[i] = kp[i].key; }
    free(kp);
}

/* ===================================================================================== */
/* a4-a6: bsa_rowReordering_gpu (src/rowReordering.cu:1027-1095).                          */
/* The concurrent cluster CTAs with hand-over-hand row mutexes (:325-432) are equivalent   */
/* to this sequential sweep: a child cluster can never overtake its parent, so each        */
/* cluster sees exactly the rows still unassigned after all earlier clusters passed.       */
/*   - rows in stable ascending dispersion order (:1055-1062)                             */
/*   - leading rows with dispersion 0 (empty) -> cluster 0 (:939-949)                      */
/*   - cluster c starts at the first row its parent rejected (:399-423); representative =  */
/*     that row's encoding; every later unassigned row joins iff sim > alpha, and then     */
/*     rep += enc (:393-395)                                                               */
/*   - permutation = stable sort of positions by cluster id (:986-995)                     */
/*   - numClusters = sortedIds[indices[M-1]] + (any empty row) (:996 -- the index is       */
/*     applied to the ALREADY SORTED key array; reproduced verbatim as clusters_compat)    */
/*   - strip leading empty rows (:1081-1090)                                               */
/* ===================================================================================== */
void oracle_row_reordering(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                           float alpha, uint32_t block_size, int exact,
                           uint32_t* perm_out, uint32_t* num_out,
                           int* clusters_compat, int* clusters_true) {
    const uint32_t nb = (uint32_t)ceilf((float)N / (float)block_size);
    const uint32_t bd = oracle_clustering_blockdim(nb);
    uint32_t* enc = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)M * nb + 4);
    uint32_t* disp = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t* asc = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t* cid = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t* rep = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)nb + 1));
    oracle_dispersion(M, N, row_offsets, col_indices, block_size, enc, disp);
    stable_argsort_u32(disp, M, asc, NULL);
    memset(cid, 0xFF, sizeof(uint32_t) * ((size_t)M + 1));   /* ORACLE_NULL_VALUE everywhere */

    uint32_t zero_row_idx = 0;
    while (zero_row_idx < M && disp[asc[zero_row_idx]] == 0) { cid[zero_row_idx] = 0; zero_row_idx++; }

    uint32_t cluster = 1, start = zero_row_idx;
    while (start < M) {
        cid[start] = cluster;
        memcpy(rep, enc + (size_t)asc[start] * nb, sizeof(uint32_t) * nb);
        uint32_t next_start = ORACLE_NULL_VALUE;
        for (uint32_t idx = start + 1; idx < M; ++idx) {
            if (cid[idx] != ORACLE_NULL_VALUE) continue;
            const uint32_t* cmp = enc + (size_t)asc[idx] * nb;
            const float sim = oracle_similarity(rep, cmp, nb, bd, exact);
            if (sim > alpha) {
                cid[idx] = cluster;
                for (uint32_t b = 0; b < nb; ++b) rep[b] += cmp[b];
            } else if (next_start == ORACLE_NULL_VALUE) {
                next_start = idx;
            }
        }
        if (next_start == ORACLE_NULL_VALUE) break;
        start = next_start;
        cluster++;
    }

    uint32_t* indices = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t* sorted = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    stable_argsort_u32(cid, M, indices, sorted);
    uint32_t* perm = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    for (uint32_t i = 0; i < M; ++i) perm[i] = asc[indices[i]];
    if (clusters_compat) *clusters_compat = M ? (int)sorted[indices[M - 1]] + (zero_row_idx != 0) : 0;
    if (clusters_true) *clusters_true = (zero_row_idx < M) ? (int)cluster : 0;

    uint32_t first = 0;
    while (first < M && row_offsets[perm[first] + 1] - row_offsets[perm[first]] == 0) ++first;
    *num_out = M - first;
    memcpy(perm_out, perm + first, sizeof(uint32_t) * (M - first));
    free(enc); free(disp); free(asc); free(cid); free(rep); free(indices); free(sorted); free(perm);
}

/* ===================================================================================== */
/* a4-a6 at graph scale: the SAME sequential semantics as oracle_row_reordering, restated    */
/* on sparse encodings so that it finishes on 10^6-row inputs (the dense rows x nb matrix    */
/* of the reference, src/rowReordering.cu:1040-1043, is 25 GB at 2^20 rows).                 */
/*   - similarity: every float operation of calculate_similarity_norm_weighted_jaccard        */
/*     (:235-293) in the same order -- thread t sums the terms i = t, t+bd, ... ascending;    */
/*     zero terms add +0.0f exactly, so only threads that own a non-zero block of the          */
/*     candidate are recomputed, the others keep the representative-only partial sums          */
/*   - candidates: a row sharing no kept block with the representative has min-sum 0, i.e.     */
/*     sim = 0 <= alpha (alpha >= 0), and  sim <= (nnz of the row in shared kept blocks) /     */
/*     (nnz of the row in kept blocks).  Every row is therefore filed only under a subset      */
/*     of its blocks chosen so that the blocks left out hold less than (alpha - 1e-3) of its   */
/*     nnz (use_filter != 0; the most popular blocks are left out first); a representative     */
/*     that contains none of the filed blocks cannot be joined by the row.  The sweep of a     */
/*     cluster visits, in position order, the unassigned rows filed under a block of the       */
/*     representative; a block that enters the representative at a join contributes its rows   */
/*     behind the join position only (earlier ones were passed while their similarity was 0). */
/*   - rows whose (possibly lossy) sum of squares is 0 join exactly the representatives whose  */
/*     sum is 0 too (:258-263): they are filed under one pseudo block.                         */
/* ===================================================================================== */
typedef struct { uint32_t* a; size_t n, cap; } u32heap;
static void heap_push(u32heap* h, uint32_t v) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->a = (uint32_t*)realloc(h->a, h->cap * sizeof(uint32_t)); }
    size_t i = h->n++;
    while (i > 0) {
        const size_t p = (i - 1) / 2;
        if (h->a[p] <= v) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = v;
}
static uint32_t heap_pop(u32heap* h) {
    const uint32_t top = h->a[0];
    const uint32_t v = h->a[--h->n];
    size_t i = 0;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && h->a[c + 1] < h->a[c]) ++c;
        if (h->a[c] >= v) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n) h->a[i] = v;
    return top;
}
typedef struct { uint32_t pop, blk, cnt; } pbc_t;
static int cmp_pbc_desc(const void* a, const void* b) {
    const pbc_t* x = (const pbc_t*)a; const pbc_t* y = (const pbc_t*)b;
    if (x->pop != y->pop) return x->pop > y->pop ? -1 : 1;
    return x->blk < y->blk ? -1 : (x->blk > y->blk);
}
/* shared-memory tree of include/cudaUtil.cuh:37-43 over per-warp values */
static float warp_tree_f32(const float* wv, uint32_t nw, uint32_t first_stride) {
    float shm[64];
    for (uint32_t i = 0; i < 64; ++i) shm[i] = i < nw ? wv[i] : 0.0f;
    for (uint32_t s = first_stride; s >= 1; s >>= 1)
        for (uint32_t w = 0; w < s; ++w) shm[w] = shm[w] + shm[w + s];
    return shm[0];
}
static float butterfly32(const float* in) {
    float v[32], t[32];
    memcpy(v, in, sizeof v);
    for (uint32_t x = 1; x < 32; x <<= 1) {
        for (uint32_t l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ x];
        memcpy(v, t, sizeof v);
    }
    return v[0];
}

void oracle_row_reordering_indexed(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                                   float alpha, uint32_t block_size, int exact, int use_filter,
                                   uint32_t* perm_out, uint32_t* num_out, int* clusters_compat, int* clusters_true,
                                   uint32_t* cluster_ids_by_pos, uint64_t* stats /* [4]: evaluations, joins, filed entries, runs */) {
    const uint32_t nb = (uint32_t)ceilf((float)N / (float)block_size);
    const uint32_t bd = oracle_clustering_blockdim(nb);
    const uint32_t nw = bd / 32;
    uint32_t first_stride = nw / 2;
    if (exact) { uint32_t p = 1; while (p < nw) p <<= 1; first_stride = p / 2; }
    uint32_t kept_mask = 0;
    {   /* which warps the tree keeps (every warp at most once) */
        uint32_t contrib[64];
        for (uint32_t w = 0; w < 64; ++w) contrib[w] = w < nw ? (1u << w) : 0u;
        for (uint32_t s = first_stride; s >= 1; s >>= 1)
            for (uint32_t w = 0; w < s; ++w) contrib[w] |= contrib[w + s];
        kept_mask = contrib[0];
    }
#define KEPT(blk) ((kept_mask >> (((blk) % bd) >> 5)) & 1u)
    const uint32_t nnz = row_offsets[M];
    /* ---- sparse encodings (:49-93) ---- */
    uint32_t* enc_ptr = (uint32_t*)calloc((size_t)M + 1, sizeof(uint32_t));
    uint32_t* enc_blk = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)nnz + 1));
    uint32_t* enc_cnt = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)nnz + 1));
    uint32_t* disp = (uint32_t*)calloc((size_t)M + 1, sizeof(uint32_t));
    uint32_t* row_sq = (uint32_t*)calloc((size_t)M + 1, sizeof(uint32_t));
    uint32_t* pop = (uint32_t*)calloc((size_t)nb + 2, sizeof(uint32_t));
    uint32_t* tmpb = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)N + 1));
    uint32_t runs = 0;
    for (uint32_t r = 0; r < M; ++r) {
        const uint32_t b = row_offsets[r], e = row_offsets[r + 1];
        enc_ptr[r] = runs;
        if (e == b) continue;
        for (uint32_t k = b; k < e; ++k) tmpb[k - b] = col_indices[k] / block_size;
        qsort(tmpb, e - b, sizeof(uint32_t), cmp_u32_fwd);
        uint32_t res = 0, sq = 0;
        const uint32_t first = runs;
        for (uint32_t k = 0; k < e - b;) {
            uint32_t k2 = k;
            while (k2 < e - b && tmpb[k2] == tmpb[k]) ++k2;
            enc_blk[runs] = tmpb[k]; enc_cnt[runs] = k2 - k; ++runs;
            res += block_size - (k2 - k);
            if (KEPT(tmpb[k])) sq += (k2 - k) * (k2 - k);
            pop[tmpb[k]]++;
            k = k2;
        }
        disp[r] = res + (e - b) * (runs - first);
        row_sq[r] = sq;
    }
    enc_ptr[M] = runs;
    free(tmpb);
    uint32_t* asc = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t* cid = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    stable_argsort_u32(disp, M, asc, NULL);
    memset(cid, 0xFF, sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t zero_row_idx = 0;
    while (zero_row_idx < M && disp[asc[zero_row_idx]] == 0) { cid[zero_row_idx] = 0; zero_row_idx++; }

    /* ---- the rows filed per block (pseudo block nb: rows whose sum of squares is 0) ---- */
    const float bound = alpha - 1e-3f;
    uint8_t* filed = (uint8_t*)calloc((size_t)runs + 1, 1);
    uint32_t* inv_ptr = (uint32_t*)calloc((size_t)nb + 3, sizeof(uint32_t));
    pbc_t* srt = (pbc_t*)malloc(sizeof(pbc_t) * ((size_t)nb + 1));
    uint64_t filed_total = 0;
    for (uint32_t p = zero_row_idx; p < M; ++p) {
        const uint32_t r = asc[p], b = enc_ptr[r], e = enc_ptr[r + 1];
        if (row_sq[r] == 0) { inv_ptr[nb + 1]++; filed_total++; continue; }
        uint32_t n = 0, tot = 0;
        for (uint32_t j = b; j < e; ++j)
            if (KEPT(enc_blk[j])) { srt[n].pop = pop[enc_blk[j]]; srt[n].blk = j; srt[n].cnt = enc_cnt[j]; ++n; tot += enc_cnt[j]; }
        if (use_filter && bound > 0.0f) qsort(srt, n, sizeof(pbc_t), cmp_pbc_desc);
        uint32_t left_out = 0;
        for (uint32_t k = 0; k < n; ++k) {
            if (use_filter && bound > 0.0f && (float)(left_out + srt[k].cnt) < bound * (float)tot) { left_out += srt[k].cnt; continue; }
            filed[srt[k].blk] = 1;
            inv_ptr[enc_blk[srt[k].blk] + 1]++;
            filed_total++;
        }
    }
    free(srt);
    for (uint32_t b = 0; b <= nb; ++b) inv_ptr[b + 1] += inv_ptr[b];
    uint32_t* inv = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)filed_total + 1));
    uint32_t* fillp = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)nb + 2));
    memcpy(fillp, inv_ptr, sizeof(uint32_t) * ((size_t)nb + 1));
    for (uint32_t p = zero_row_idx; p < M; ++p) {          /* ascending positions -> every list is sorted */
        const uint32_t r = asc[p];
        if (row_sq[r] == 0) { inv[fillp[nb]++] = p; continue; }
        for (uint32_t j = enc_ptr[r]; j < enc_ptr[r + 1]; ++j)
            if (filed[j]) inv[fillp[enc_blk[j]]++] = p;
    }
    free(fillp); free(filed);

    /* ---- the sweep ---- */
    uint32_t* rep = (uint32_t*)calloc((size_t)nb + 1, sizeof(uint32_t));
    float* repn = (float*)calloc((size_t)nb + 1, sizeof(float));
    uint32_t* candv = (uint32_t*)calloc((size_t)nb + 1, sizeof(uint32_t));
    uint32_t* rep_blocks = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)nb + 1));
    uint32_t* merged = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)nb + 1));
    uint32_t* stamp = (uint32_t*)calloc((size_t)M + 1, sizeof(uint32_t));
    float part_max[1024], warp_max[32];
    uint8_t thr_touched[1024];
    memset(thr_touched, 0, sizeof thr_touched);
    u32heap heap = {0, 0, 0};
    uint64_t n_eval = 0, n_join = 0;
    uint32_t cluster = 0, start = zero_row_idx;
    while (start < M) {
        ++cluster;
        cid[start] = cluster;
        uint32_t n_rep_blocks = 0, s_rep = 0;
        float n_rep = 0.0f;
        heap.n = 0;
        uint32_t last_pos = start;
        /* absorb the rows of `pos` into the representative and refresh everything derived from it */
#define ABSORB(pos)                                                                                        \
        do {                                                                                               \
            const uint32_t r_ = asc[(pos)], b_ = enc_ptr[r_], e_ = enc_ptr[r_ + 1];                        \
            uint32_t i_ = 0, j_ = b_, m_ = 0;                                                              \
            while (i_ < n_rep_blocks || j_ < e_) {                                                         \
                if (j_ >= e_ || (i_ < n_rep_blocks && rep_blocks[i_] < enc_blk[j_])) merged[m_++] = rep_blocks[i_++]; \
                else if (i_ >= n_rep_blocks || enc_blk[j_] < rep_blocks[i_]) {                             \
                    const uint32_t nbk_ = enc_blk[j_];                                                     \
                    merged[m_++] = nbk_;                                                                   \
                    if (KEPT(nbk_)) {   /* a block new to the representative: its rows behind this position */ \
                        uint32_t lo_ = inv_ptr[nbk_], hi_ = inv_ptr[nbk_ + 1];                             \
                        while (lo_ < hi_) { const uint32_t mid_ = (lo_ + hi_) / 2; if (inv[mid_] <= (pos)) lo_ = mid_ + 1; else hi_ = mid_; } \
                        for (uint32_t q_ = lo_; q_ < inv_ptr[nbk_ + 1]; ++q_) {                            \
                            const uint32_t cp_ = inv[q_];                                                  \
                            if (cid[cp_] == ORACLE_NULL_VALUE && stamp[cp_] != cluster) { stamp[cp_] = cluster; heap_push(&heap, cp_); } \
                        }                                                                                  \
                    }                                                                                      \
                    ++j_;                                                                                  \
                } else { merged[m_++] = rep_blocks[i_++]; ++j_; }                                          \
            }                                                                                              \
            for (uint32_t j2_ = b_; j2_ < e_; ++j2_) rep[enc_blk[j2_]] += enc_cnt[j2_];                    \
            memcpy(rep_blocks, merged, sizeof(uint32_t) * m_);                                             \
            n_rep_blocks = m_;                                                                             \
            s_rep = 0;                                                                                     \
            for (uint32_t k_ = 0; k_ < n_rep_blocks; ++k_)                                                 \
                if (KEPT(rep_blocks[k_])) s_rep += rep[rep_blocks[k_]] * rep[rep_blocks[k_]];              \
            n_rep = sqrtf((float)s_rep);                                                                   \
            for (uint32_t t_ = 0; t_ < bd; ++t_) part_max[t_] = 0.0f;                                      \
            for (uint32_t k_ = 0; k_ < n_rep_blocks; ++k_) {   /* ascending i -> ascending inside every thread */ \
                const uint32_t i2_ = rep_blocks[k_];                                                       \
                repn[i2_] = (float)rep[i2_] / n_rep;                                                       \
                part_max[i2_ % bd] = part_max[i2_ % bd] + repn[i2_];                                       \
            }                                                                                              \
            for (uint32_t w_ = 0; w_ < nw; ++w_) warp_max[w_] = butterfly32(part_max + 32 * w_);           \
        } while (0)
        ABSORB(start);
        if (s_rep == 0) {
            /* a representative without a kept block: exactly the rows of the same kind can join (:258-260) */
            uint32_t lo = inv_ptr[nb], hi = inv_ptr[nb + 1];
            while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (inv[mid] <= start) lo = mid + 1; else hi = mid; }
            for (uint32_t q = lo; q < inv_ptr[nb + 1]; ++q)
                if (cid[inv[q]] == ORACLE_NULL_VALUE && stamp[inv[q]] != cluster) { stamp[inv[q]] = cluster; heap_push(&heap, inv[q]); }
        }
        while (heap.n) {
            const uint32_t p = heap_pop(&heap);
            if (p <= last_pos || cid[p] != ORACLE_NULL_VALUE) continue;
            last_pos = p;
            const uint32_t r = asc[p], b = enc_ptr[r], e = enc_ptr[r + 1];
            const uint32_t s_cmp = row_sq[r];
            ++n_eval;
            float sim;
            if (s_rep == 0 && s_cmp == 0) sim = 1.0f;
            else if (s_rep == 0 || s_cmp == 0) sim = 0.0f;
            else {
                const float n_cmp = sqrtf((float)s_cmp);
                float wmin[32], wmax[32];
                uint32_t touched_warps = 0;
                for (uint32_t j = b; j < e; ++j) { candv[enc_blk[j]] = enc_cnt[j]; thr_touched[enc_blk[j] % bd] = 1; touched_warps |= 1u << ((enc_blk[j] % bd) >> 5); }
                for (uint32_t w = 0; w < nw; ++w) {
                    if (!((touched_warps >> w) & 1u)) { wmin[w] = 0.0f; wmax[w] = warp_max[w]; continue; }
                    float lmin[32], lmax[32];
                    for (uint32_t l = 0; l < 32; ++l) {
                        const uint32_t t = 32 * w + l;
                        if (!thr_touched[t]) { lmin[l] = 0.0f; lmax[l] = part_max[t]; continue; }
                        float mn = 0.0f, mx = 0.0f;
                        for (uint32_t i = t; i < nb; i += bd) {
                            const float a = repn[i];
                            const float c = (float)candv[i] / n_cmp;
                            mn = mn + fminf(a, c);
                            mx = mx + fmaxf(a, c);
                        }
                        lmin[l] = mn; lmax[l] = mx;
                    }
                    wmin[w] = butterfly32(lmin);
                    wmax[w] = butterfly32(lmax);
                }
                for (uint32_t j = b; j < e; ++j) { candv[enc_blk[j]] = 0; thr_touched[enc_blk[j] % bd] = 0; }
                sim = warp_tree_f32(wmin, nw, first_stride) / warp_tree_f32(wmax, nw, first_stride);
            }
            if (sim > alpha) {
                cid[p] = cluster;
                ++n_join;
                ABSORB(p);
            }
        }
        for (uint32_t k = 0; k < n_rep_blocks; ++k) { rep[rep_blocks[k]] = 0; repn[rep_blocks[k]] = 0.0f; }
        while (start < M && cid[start] != ORACLE_NULL_VALUE) ++start;
    }
#undef ABSORB
#undef KEPT
    if (stats) { stats[0] = n_eval; stats[1] = n_join; stats[2] = filed_total; stats[3] = runs; }
    if (cluster_ids_by_pos) memcpy(cluster_ids_by_pos, cid, sizeof(uint32_t) * M);

    uint32_t* indices = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    uint32_t* sorted = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    stable_argsort_u32(cid, M, indices, sorted);
    uint32_t* perm = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)M + 1));
    for (uint32_t i = 0; i < M; ++i) perm[i] = asc[indices[i]];
    if (clusters_compat) *clusters_compat = M ? (int)sorted[indices[M - 1]] + (zero_row_idx != 0) : 0;
    if (clusters_true) *clusters_true = (zero_row_idx < M) ? (int)cluster : 0;
    uint32_t first = 0;
    while (first < M && row_offsets[perm[first] + 1] - row_offsets[perm[first]] == 0) ++first;
    *num_out = M - first;
    memcpy(perm_out, perm + first, sizeof(uint32_t) * (M - first));
    free(enc_ptr); free(enc_blk); free(enc_cnt); free(disp); free(row_sq); free(pop); free(asc); free(cid);
    free(inv_ptr); free(inv); free(rep); free(repn); free(candv); free(rep_blocks); free(merged); free(stamp);
    free(heap.a); free(indices); free(sorted); free(perm);
}

/* ===================================================================================== */
/* a8: colReordering_cpu (src/colReordering.cu:274-404) + analysisDescendingOrderColSegment */
/* (:244-271).  Per 16-row panel of the reordered rows:                                    */
/*   - nnz count per column (:300-312), keep non-empty columns in ascending id (:318-331)  */
/*   - stable sort by count DESCENDING (:333-336)  -> ties keep ascending column id        */
/*   - pad to a multiple of 16 with sentinel column = N, count 0 (:338-343)                */
/*   - threshold = (UIN)ceil(delta * 256); every full 16-column block whose nnz >= thr      */
/*     adds 16 to the dense count (:249-261); counts are sorted so dense blocks are a      */
/*     prefix; dense = first numDense columns, sparse = the rest INCLUDING sentinel pads    */
/*   - sparse value count = sum of counts over the sparse columns (:350-353)               */
/*   - three exclusive scans (:360-378)                                                    */
/* ===================================================================================== */
static int cmp_u32(const void* a, const void* b) {
    const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : (x > y);
}

void oracle_col_reordering(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                           const uint32_t* rows, uint32_t num_rows, float delta, oracle_colreorder* out) {
    (void)M;
    const uint32_t panels = (uint32_t)ceilf((float)num_rows / (float)ORACLE_ROW_PANEL_SIZE); /* BSMR.cpp:48 */
    const uint32_t thr = (uint32_t)ceilf(delta * (float)(ORACLE_ROW_PANEL_SIZE * ORACLE_BLOCK_COL_SIZE));
    memset(out, 0, sizeof(*out));
    out->num_row_panels = panels;
    out->dense_col_offsets = (uint32_t*)calloc((size_t)panels + 1, sizeof(uint32_t));
    out->sparse_col_offsets = (uint32_t*)calloc((size_t)panels + 1, sizeof(uint32_t));
    out->sparse_value_offsets = (uint32_t*)calloc((size_t)panels + 1, sizeof(uint32_t));
    uint32_t** panel_cols = (uint32_t**)calloc(panels ? panels : 1, sizeof(uint32_t*));
    uint32_t* panel_ncols = (uint32_t*)calloc(panels ? panels : 1, sizeof(uint32_t));
    uint32_t* panel_ndense = (uint32_t*)calloc(panels ? panels : 1, sizeof(uint32_t));

    for (uint32_t p = 0; p < panels; ++p) {
        const uint32_t r0 = p * ORACLE_ROW_PANEL_SIZE;
        const uint32_t r1 = r0 + ORACLE_ROW_PANEL_SIZE < num_rows ? r0 + ORACLE_ROW_PANEL_SIZE : num_rows;
        size_t tot = 0;
        for (uint32_t i = r0; i < r1; ++i) tot += row_offsets[rows[i] + 1] - row_offsets[rows[i]];
        uint32_t* all = (uint32_t*)malloc(sizeof(uint32_t) * (tot ? tot : 1));
        size_t n = 0;
        for (uint32_t i = r0; i < r1; ++i)
            for (uint32_t k = row_offsets[rows[i]]; k < row_offsets[rows[i] + 1]; ++k) all[n++] = col_indices[k];
        qsort(all, n, sizeof(uint32_t), cmp_u32);
        /* unique columns ascending + counts */
        uint32_t* ucol = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
        uint32_t* ucnt = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
        uint32_t nu = 0;
        for (size_t i = 0; i < n; ++i) {
            if (nu && ucol[nu - 1] == all[i]) ucnt[nu - 1]++;
            else { ucol[nu] = all[i]; ucnt[nu] = 1; nu++; }
        }
        /* stable descending sort by count: bucket passes from high count to low */
        uint32_t padded = nu % ORACLE_BLOCK_COL_SIZE ? nu + ORACLE_BLOCK_COL_SIZE - nu % ORACLE_BLOCK_COL_SIZE : nu;
        uint32_t* scol = (uint32_t*)malloc(sizeof(uint32_t) * (padded ? padded : 1));
        uint32_t* scnt = (uint32_t*)malloc(sizeof(uint32_t) * (padded ? padded : 1));
        uint32_t maxc = 0;
        for (uint32_t i = 0; i < nu; ++i) if (ucnt[i] > maxc) maxc = ucnt[i];
        uint32_t w = 0;
        for (uint32_t c = maxc; c >= 1; --c)
            for (uint32_t i = 0; i < nu; ++i)
                if (ucnt[i] == c) { scol[w] = ucol[i]; scnt[w] = c; w++; }
        for (; w < padded; ++w) { scol[w] = N; scnt[w] = 0; }
        /* analysisDescendingOrderColSegment */
        uint32_t seg = 0, ndense = 0;
        while (seg + ORACLE_BLOCK_COL_SIZE <= padded) {
            uint32_t s = 0;
            for (uint32_t i = 0; i < ORACLE_BLOCK_COL_SIZE; ++i) s += scnt[seg + i];
            if (s >= thr) ndense += ORACLE_BLOCK_COL_SIZE;
            seg += ORACLE_BLOCK_COL_SIZE;
        }
        while (seg < padded && scnt[seg] > 0) ++seg;
        const uint32_t nsparse = seg - ndense;
        uint32_t sparse_data = 0;
        for (uint32_t i = ndense; i < ndense + nsparse; ++i) sparse_data += scnt[i];
        panel_cols[p] = scol; panel_ncols[p] = padded; panel_ndense[p] = ndense;
        out->dense_col_offsets[p + 1] = ndense;
        out->sparse_col_offsets[p + 1] = nsparse;
        out->sparse_value_offsets[p + 1] = sparse_data;
        free(all); free(ucol); free(ucnt); free(scnt);
    }
    for (uint32_t p = 0; p < panels; ++p) {
        out->dense_col_offsets[p + 1] += out->dense_col_offsets[p];
        out->sparse_col_offsets[p + 1] += out->sparse_col_offsets[p];
        out->sparse_value_offsets[p + 1] += out->sparse_value_offsets[p];
    }
    out->n_dense_cols = out->dense_col_offsets[panels];
    out->n_sparse_cols = out->sparse_col_offsets[panels];
    out->dense_cols = (uint32_t*)malloc(sizeof(uint32_t) * (out->n_dense_cols ? out->n_dense_cols : 1));
    out->sparse_cols = (uint32_t*)malloc(sizeof(uint32_t) * (out->n_sparse_cols ? out->n_sparse_cols : 1));
    for (uint32_t p = 0; p < panels; ++p) {
        /* dense = [0, ndense), sparse = [ndense, end of the padded list) (:386-399) */
        memcpy(out->dense_cols + out->dense_col_offsets[p], panel_cols[p], sizeof(uint32_t) * panel_ndense[p]);
        memcpy(out->sparse_cols + out->sparse_col_offsets[p], panel_cols[p] + panel_ndense[p],
               sizeof(uint32_t) * (panel_ncols[p] - panel_ndense[p]));
        free(panel_cols[p]);
    }
    free(panel_cols); free(panel_ncols); free(panel_ndense);
}

void oracle_free_colreorder(oracle_colreorder* r) {
    free(r->dense_cols); free(r->dense_col_offsets); free(r->sparse_cols);
    free(r->sparse_col_offsets); free(r->sparse_value_offsets);
    memset(r, 0, sizeof(*r));
}

/* ===================================================================================== */
/* a9: RPHM (src/BSMR.cpp:83-265), data part only (the per-CTA work lists are launch       */
/* details of the reference kernels).                                                      */
/*   blockOffsets[p+1] = blockOffsets[p] + ceil(#denseCols_p / 16)            (:125-136)   */
/*   blockValues[(blockOffsets[p]+cb)*256 + r*16 + c] = CSR index of (row r of panel p,     */
/*       dense column cb*16+c) or NULL_VALUE                                   (:143-174)   */
/*   residual triplets ordered by (panel, residual column order, row-in-panel) (:177-219)  */
/* ===================================================================================== */
void oracle_build_rphm(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                       const uint32_t* rows, uint32_t num_rows, const oracle_colreorder* cr, oracle_rphm* out) {
    (void)M; (void)N;
    const uint32_t panels = cr->num_row_panels;
    memset(out, 0, sizeof(*out));
    out->num_row_panels = panels;
    out->block_offsets = (uint32_t*)calloc((size_t)panels + 1, sizeof(uint32_t));
    for (uint32_t p = 0; p < panels; ++p) {
        const uint32_t nc = cr->dense_col_offsets[p + 1] - cr->dense_col_offsets[p];
        out->block_offsets[p + 1] = out->block_offsets[p] + (nc + ORACLE_BLOCK_COL_SIZE - 1) / ORACLE_BLOCK_COL_SIZE;
    }
    out->n_block_values = (size_t)out->block_offsets[panels] * 256u;
    out->block_values = (uint32_t*)malloc(sizeof(uint32_t) * (out->n_block_values ? out->n_block_values : 1));
    for (size_t i = 0; i < out->n_block_values; ++i) out->block_values[i] = ORACLE_NULL_VALUE;
    out->n_sparse = cr->sparse_value_offsets[panels];
    const size_t ns = out->n_sparse ? out->n_sparse : 1;
    out->sparse_values = (uint32_t*)malloc(sizeof(uint32_t) * ns);
    out->sparse_relative_rows = (uint32_t*)malloc(sizeof(uint32_t) * ns);
    out->sparse_col_indices = (uint32_t*)malloc(sizeof(uint32_t) * ns);

    for (uint32_t p = 0; p < panels; ++p) {
        const uint32_t r0 = p * ORACLE_ROW_PANEL_SIZE;
        const uint32_t r1 = r0 + ORACLE_ROW_PANEL_SIZE < num_rows ? r0 + ORACLE_ROW_PANEL_SIZE : num_rows;
        /* dense part */
        const uint32_t dc0 = cr->dense_col_offsets[p], dc1 = cr->dense_col_offsets[p + 1];
        for (uint32_t ri = r0; ri < r1; ++ri) {
            const uint32_t row = rows[ri];
            for (uint32_t j = dc0; j < dc1; ++j) {
                const uint32_t col = cr->dense_cols[j];
                for (uint32_t k = row_offsets[row]; k < row_offsets[row + 1]; ++k) {
                    if (col_indices[k] == col) {
                        const uint32_t cnt = j - dc0;
                        const size_t at = ((size_t)out->block_offsets[p] + cnt / 16u) * 256u + (ri - r0) * 16u + cnt % 16u;
                        out->block_values[at] = k;
                    }
                }
            }
        }
        /* residual part */
        uint32_t w = cr->sparse_value_offsets[p];
        for (uint32_t j = cr->sparse_col_offsets[p]; j < cr->sparse_col_offsets[p + 1]; ++j) {
            const uint32_t col = cr->sparse_cols[j];
            for (uint32_t ri = r0; ri < r1; ++ri) {
                const uint32_t row = rows[ri];
                for (uint32_t k = row_offsets[row]; k < row_offsets[row + 1]; ++k) {
                    if (col_indices[k] == col) {
                        out->sparse_relative_rows[w] = ri - r0;
                        out->sparse_values[w] = k;
                        out->sparse_col_indices[w] = col;
                        ++w;
                    }
                }
            }
        }
    }
}

void oracle_free_rphm(oracle_rphm* r) {
    free(r->block_offsets); free(r->block_values); free(r->sparse_values);
    free(r->sparse_relative_rows); free(r->sparse_col_indices);
    memset(r, 0, sizeof(*r));
}
