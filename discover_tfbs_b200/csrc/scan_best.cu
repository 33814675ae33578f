// scan_best.cu — per-record best PWM window on both strands (the PWM columns of the feature matrix).
//
// north_star item (3) "writes the feature matrix consumed by the classifier": the reference joins its
// per-site feature columns in utils.py:1134-1170; the PWM columns this package adds
// (`<name>_fwd_max`, `<name>_rev_max`) are max over the record's windows of Biopython's
// PositionSpecificScoringMatrix.calculate (Bio/motifs/_pwm.c).  Round 1 downloaded the
// dense [M][n] score arrays and reduced them with NumPy nanmax (530 MB D2H for config C2); this kernel
// reduces on the device and writes float[S][M][2] (+ the offset of the best window) only.
//
// Arithmetic: every window is scored exactly as the oracle does — left-to-right double sum of the
// exact log-odds entries, cast to float — and the maximum is taken over the FLOAT scores (two windows
// whose double sums differ may be equal as floats; numpy.nanargmax then reports the first), so values
// and offsets are BIT-EXACT with max / first argmax over the oracle's calculate().
// One warp per (record, motif); lanes stride over the window starts; windows touching an invalid
// byte are skipped (np.nanmax semantics); a record without a valid window yields NaN / -1.
#include <cmath>

#include "common.cuh"

namespace {

constexpr int kBestThreads = 256;  // 8 warps

struct BestParams {
    const uint32_t *packed;
    const uint32_t *mask;
    const double *exact;     // [M][2][TFBS_MAX_MOTIF_LEN][4]
    const int32_t *lens;     // [M]
    const int64_t *rec_start;
    const int64_t *rec_len;
    int64_t S;
    int32_t M;
    float *best;             // [S][M][2]
    int32_t *pos;            // [S][M][2] or nullptr
};

__device__ __forceinline__ void take_better(float &v, int32_t &p, float ov, int32_t op)
{
    // larger score wins; equal scores keep the smaller offset (np.nanargmax: first maximum); p < 0 = none yet
    if (op >= 0 && (p < 0 || ov > v || (ov == v && op < p))) { v = ov; p = op; }
}

__global__ void __launch_bounds__(kBestThreads) scan_best_kernel(const BestParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kBestThreads + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * kBestThreads) >> 5;
    const int64_t total = P.S * P.M;
    for (int64_t w = warp; w < total; w += n_warps) {
        const int64_t s = w / P.M;
        const int m = (int)(w - s * P.M);
        const int L = __ldg(P.lens + m);
        const int64_t r0 = __ldg(P.rec_start + s);
        const int64_t nwin = __ldg(P.rec_len + s) - L + 1;
        const double *F = P.exact + (size_t)m * 2 * TFBS_MAX_MOTIF_LEN * 4;
        const double *R = F + TFBS_MAX_MOTIF_LEN * 4;
        const uint32_t lmask = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
        float bf = 0.f, br = 0.f;
        int32_t pf = -1, pr = -1;
        for (int64_t i = lane; i < nwin; i += 32) {
            const int64_t p = r0 + i;
            const uint32_t *mp = P.mask + (p >> 5);
            const uint32_t inv = __funnelshift_r(__ldg(mp), __ldg(mp + 1), (int)(p & 31)) & lmask;
            if (inv) continue;
            const uint32_t *wp = P.packed + (p >> 4);
            const uint32_t a0 = __ldg(wp), a1 = __ldg(wp + 1), a2 = __ldg(wp + 2);
            const int sh = (int)(p & 15) * 2;
            uint64_t bases = (uint64_t)__funnelshift_r(a0, a1, sh) | ((uint64_t)__funnelshift_r(a1, a2, sh) << 32);
            double sf = 0.0, sr = 0.0;
            for (int j = 0; j < L; j++) {
                const int c = (int)(bases & 3u);
                bases >>= 2;
                sf += __ldg(F + j * 4 + c);
                sr += __ldg(R + j * 4 + c);
            }
            take_better(bf, pf, (float)sf, (int32_t)i);
            take_better(br, pr, (float)sr, (int32_t)i);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float of = __shfl_xor_sync(0xFFFFFFFFu, bf, d), orv = __shfl_xor_sync(0xFFFFFFFFu, br, d);
            const int32_t opf = __shfl_xor_sync(0xFFFFFFFFu, pf, d), opr = __shfl_xor_sync(0xFFFFFFFFu, pr, d);
            take_better(bf, pf, of, opf);
            take_better(br, pr, orv, opr);
        }
        if (lane == 0) {
            const float qnan = __uint_as_float(0x7FC00000u);
            P.best[w * 2 + 0] = pf >= 0 ? bf : qnan;
            P.best[w * 2 + 1] = pr >= 0 ? br : qnan;
            if (P.pos) {
                P.pos[w * 2 + 0] = pf;
                P.pos[w * 2 + 1] = pr;
            }
        }
    }
}

}  // namespace

extern "C" int tfbs_scan_best(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_pwms *pwms, const int64_t *rec_start,
                              const int64_t *rec_len, int64_t S, float *best_host, int32_t *pos_host)
{
    TFBS_REQUIRE(ctx && seq && pwms && S >= 0 && (S == 0 || (rec_start && rec_len)), "tfbs_scan_best: null argument");
    TFBS_REQUIRE(seq->ctx == ctx && pwms->ctx == ctx, "tfbs_scan_best: handle from another context");
    if (S == 0) return TFBS_OK;
    for (int64_t i = 0; i < S; i++)
        TFBS_REQUIRE(rec_start[i] >= 0 && rec_len[i] >= 0 && rec_start[i] + rec_len[i] <= seq->n &&
                         rec_len[i] < ((int64_t)1 << 31),
                     "tfbs_scan_best: record %lld out of range", (long long)i);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int M = pwms->M;
    const size_t cells = (size_t)S * M * 2;
    // scratch: [rec_start S][rec_len S] int64, then best float[cells], pos int32[cells]
    int rc = tfbs_scratch_reserve(ctx->sites, (size_t)S * 16 + cells * 8);
    if (rc) return rc;
    int64_t *d_start = (int64_t *)ctx->sites.ptr, *d_len = d_start + S;
    float *d_best = (float *)(d_len + S);
    int32_t *d_pos = (int32_t *)(d_best + cells);
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_start, rec_start, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_len, rec_len, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    BestParams P;
    P.packed = seq->packed;
    P.mask = seq->mask;
    P.exact = pwms->exact;
    P.lens = pwms->lens_dev;
    P.rec_start = d_start;
    P.rec_len = d_len;
    P.S = S;
    P.M = M;
    P.best = d_best;
    P.pos = pos_host ? d_pos : nullptr;
    const int64_t warps = S * (int64_t)M;
    int64_t blocks = (warps + kBestThreads / 32 - 1) / (kBestThreads / 32);
    const int64_t max_blocks = (int64_t)ctx->sm_count * 64;  // grid-stride beyond 8 resident CTAs x 8 waves per SM
    if (blocks > max_blocks) blocks = max_blocks;
    tfbs_begin(ctx);
    scan_best_kernel<<<(unsigned)blocks, kBestThreads, 0, ctx->stream>>>(P);
    tfbs_launched(ctx);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    if (best_host)
        TFBS_CHECK_CUDA(cudaMemcpyAsync(best_host, d_best, cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (pos_host)
        TFBS_CHECK_CUDA(cudaMemcpyAsync(pos_host, d_pos, cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return TFBS_OK;
}
