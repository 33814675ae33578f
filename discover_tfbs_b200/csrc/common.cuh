// common.cuh — shared declarations for libtfbs_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../include/tfbs_cuda.h"

#define TFBS_SM_COUNT_FALLBACK 148

void tfbs_set_error(const char *fmt, ...);

#define TFBS_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            tfbs_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                      \
            return _e == cudaErrorMemoryAllocation ? TFBS_ERR_NOMEM : TFBS_ERR_CUDA;        \
        }                                                                                  \
    } while (0)

#define TFBS_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            tfbs_set_error(__VA_ARGS__); \
            return TFBS_ERR_INVALID;     \
        }                                \
    } while (0)

// Grow-only device scratch buffer.
struct tfbs_scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct tfbs_ctx {
    int device = 0;
    int sm_count = TFBS_SM_COUNT_FALLBACK;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // per-call kernel bracket
    cudaEvent_t ev_mid = nullptr;              // between the filter scan and the verify pass of tfbs_scan_hits
    cudaEvent_t ev_direct = nullptr;           // between the direct-path kernel and the table filter
    float last_scan_ms = 0.f, last_verify_ms = 0.f, last_direct_ms = 0.f;
    cudaEvent_t evt0 = nullptr, evt1 = nullptr;  // user region timer
    cudaStream_t stream_in = nullptr, stream_out = nullptr;  // copy streams of the pipelined host path
    std::vector<cudaEvent_t> pipe_events;
    unsigned long long *pinned_counts = nullptr;  // pinned host scratch for per-chunk hit counts
    float last_ms = 0.f;
    int64_t last_launches = 0;
    int64_t total_launches = 0;
    int64_t last_candidates = 0;  // candidates the last hits scan verified exactly
    tfbs_scratch dense_fwd, dense_rev;  // dense scan outputs [M][pitch]
    tfbs_scratch hits;                  // tfbs_hit[cap]
    tfbs_scratch hit_count;             // unsigned long long + work counters
    tfbs_scratch cand;                  // candidate list of the hits scan (8 B records)
    tfbs_scratch tracks;                // shape tracks [T][pitch]
    tfbs_scratch sites;                 // shape sites out / args
    tfbs_scratch misc;                  // small argument staging
    tfbs_scratch flush;                 // >L2 buffer for tfbs_flush_l2
    tfbs_scratch thr;                   // per-motif thresholds
    tfbs_scratch sort;                  // radix-sort double buffers for ordered hit output
};

struct tfbs_seq {
    tfbs_ctx *ctx = nullptr;
    int64_t n = 0;          // bases
    int64_t n_words = 0;    // packed words allocated (padded, zero-filled past n)
    int64_t n_mask = 0;     // mask words allocated (padded, all-ones past n)
    uint32_t *packed = nullptr;     // first word of base 0; 64 words before it are allocated too
    uint32_t *mask = nullptr;       // (packed: zeros, mask: all ones) so kernels may read index -1
    uint32_t *packed_base = nullptr, *mask_base = nullptr;  // what cudaMalloc returned
    void *staging = nullptr;   // device ASCII staging for tfbs_encode (host input)
    size_t staging_bytes = 0;
};

// Motif library on device.
struct tfbs_pwms {
    tfbs_ctx *ctx = nullptr;
    int32_t M = 0, Lmax = 0;
    std::vector<int32_t> lens;       // host copy
    std::vector<double> logodds;     // host copy [M][Lmax][4]
    // exact matrices for re-scoring: double[M][2][TFBS_MAX_MOTIF_LEN][4] (strand 0 fwd, 1 revcomp)
    double *exact = nullptr;
    // verify pass, fast path: column-PAIR sums double[M][2][16][16] (pair j, index code(2j) + 4 code(2j+1);
    // a missing second column adds 0.0) and the per-motif rounding guard 2^-46 * sum_j max_b |entry|
    double *exact2 = nullptr;
    double *guard_eps = nullptr;
    // dense mode: fixed-point k-mer LUTs (k = 4 for L <= 24, else 3): per motif G groups of 4^k int2
    // (fwd, rev) = rint(group sum * 2^e), per-motif scale 2^-e and -inf detection bound
    int2 *lut = nullptr;
    float *dense_scale = nullptr;    // [M]
    int32_t *dense_neg = nullptr;    // [M]
    int32_t *lut_off = nullptr;      // [M] offset in int2 units
    int32_t *glen = nullptr;         // [M] packed (K << 16) | (G << 8) | L
    int64_t lut_entries = 0;
    size_t dense_lut_bytes = 0;      // shared memory the dense kernel needs for its largest table
    int32_t Gmax = 0;
    // ---- hits mode (scan_hits.cu): motifs sorted by length and cut into blocks.  A WIDE block
    // holds 32 motifs (one per lane, 16-bit filter fields), a NARROW block 64 (two per lane, 8-bit
    // fields).  The cut depends on the thresholds, so it is made with the filter tables.
    std::vector<int32_t> order, lens_desc;  // motif indices by descending length, their lengths
    int32_t n_blocks = 0, n_blocks_max = 0; // blocks of the current plan / capacity of the device arrays
    std::vector<int32_t> blk_A, blk_B, blk_narrow, blk_count, blk_lut_off, blk_motif;  // [n_blocks] x5, [n_blocks][64]
    std::vector<double> exact_host;  // host copy of `exact`
    size_t hits_lut_bytes = 0;       // largest block table (A * 32 KB + B * 8 KB)
    int64_t hq_lut_words = 0;        // total uint32 entries of the packed deficit tables
    uint32_t *hq_lut_dev = nullptr, *hq_kinit_dev = nullptr;
    int32_t *blk_AB_dev = nullptr, *blk_lut_off_dev = nullptr, *blk_motif_dev = nullptr, *lens_dev = nullptr;
    // direct path of the hits scan: the motifs of at most `direct_len` columns (the tail of `order`) are not
    // filtered through the shared-memory tables at all when thresholds are sparse: every L-mer that
    // reaches the threshold is enumerated on the host into a 12-mer hash (global memory) behind a 10-mer
    // bitmap (128 KB, shared memory); see scan_direct_kernel
    int32_t n_direct = 0, direct_len = 0;
    int64_t direct_entries = 0;
    uint32_t *direct_bitmap_dev = nullptr;
    uint2 *direct_hash_dev = nullptr;
    size_t direct_hash_slots_dev = 0;   // slots allocated on the device
    uint32_t direct_hash_mask = 0, direct_hash_shift = 0;
    std::vector<float> hq_thr;       // thresholds the cached tables were built for
    bool hq_valid = false;
};

struct tfbs_shapes {
    tfbs_ctx *ctx = nullptr;
    int32_t T = 0;
    std::vector<int32_t> kinds;
    // tables re-indexed by the little-endian pentamer index used by the kernels:
    // float[T][2][1024] on device
    float *tables = nullptr;
    int32_t *kinds_dev = nullptr;
    // packed int16 centi-unit planes (only when every table value is an exact 2-decimal number)
    uint32_t *c16_words = nullptr;
    std::vector<int32_t> c16_kind, c16_row_a, c16_row_b;
};

int tfbs_scratch_reserve(tfbs_scratch &s, size_t bytes);
void tfbs_hits_group_plan(int L, int *A_out, int *B_out);
void tfbs_hits_order_and_capacity(tfbs_pwms *p, const int32_t *lens);
void tfbs_dense_quantise(const double *F, const double *R, int L, int *K_out, int *G_out, int *e_out,
                         int32_t *neg_thr, std::vector<int2> &lut);
// Exact scoring matrices of motif m: F = log-odds as given, R = reverse complement (columns reversed,
// bases complemented), both [L][4] in a [TFBS_MAX_MOTIF_LEN][4] slab.
static inline void tfbs_fill_exact(const double *logodds, int Lmax, int m, int L, double *F, double *R)
{
    for (int j = 0; j < L; j++)
        for (int b = 0; b < 4; b++) {
            F[j * 4 + b] = logodds[((size_t)m * Lmax + j) * 4 + b];
            R[j * 4 + b] = logodds[((size_t)m * Lmax + (L - 1 - j)) * 4 + (3 - b)];
        }
}

// One block of the hits-mode plan: A 4-column + B 3-column groups, narrow (64 motifs, 8-bit fields)
// or wide (32 motifs, 16-bit fields), motifs [first, first + count) of the length-sorted order.
struct tfbs_hits_block { int A, B, narrow /* 0 wide, 1 narrow, 2 sticky narrow */, first, count; };
#define TFBS_BLOCK_SLOTS 64          // motif slots per hits-mode block (32 used by a wide block)
#define TFBS_NARROW_MAX_GROUPS 9     // narrow blocks: 8-bit fields leave (128 - 2G) / (G - 1) quantisation steps
#define TFBS_STICKY_MIN_GROUPS 5     // sticky narrow blocks (flag bits masked out of every add) from 5 groups on
int tfbs_encode_range(tfbs_ctx *ctx, const void *ascii_dev, int64_t valid, tfbs_seq *s, int64_t base,
                      int64_t cover, cudaStream_t stream);

// RAII-free timing helpers: bracket the kernels of one API call.
static inline void tfbs_begin(tfbs_ctx *ctx)
{
    ctx->last_launches = 0;
    cudaEventRecord(ctx->ev0, ctx->stream);
}
static inline void tfbs_launched(tfbs_ctx *ctx, int n = 1)
{
    ctx->last_launches += n;
    ctx->total_launches += n;
}
static inline int tfbs_end(tfbs_ctx *ctx)
{
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(ctx->ev1);
    if (e != cudaSuccess) {
        tfbs_set_error("kernel execution failed: %s", cudaGetErrorString(e));
        return TFBS_ERR_CUDA;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        tfbs_set_error("kernel launch failed: %s", cudaGetErrorString(e));
        return TFBS_ERR_CUDA;
    }
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return TFBS_OK;
}

// ---- device helpers -----------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_words_guard(const uint32_t *__restrict__ p, int64_t idx,
                                                   int64_t n_words, uint32_t fill)
{
    return (idx >= 0 && idx < n_words) ? __ldg(p + idx) : fill;
}
#endif
