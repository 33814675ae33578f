// scan.cu — (2) PWM log-odds sliding-window scan on both strands.
//
// Replaces Biopython PositionSpecificScoringMatrix.calculate/.search (Bio/motifs/matrix.py,
// Bio/motifs/_pwm.c) — the scoring the north star names — and generalises the
// reference's exact-match BLASTn candidate stage (utils.py:340-398, process_blast.sh:32).
//
// Arithmetic: the motif matrix and its reverse complement are folded on the host into k-mer
// look-up tables (k = 4 for L <= 24, else 3; each entry the double sum of its <= k log-odds values,
// fwd and rev packed in 8 bytes) so a window of length L costs ceil(L/k) shared-memory loads for
// both strands.  The LUT index is sliced straight out of the 2-bit packed sequence held in registers.
//   dense mode (this file): FIXED-POINT int32 accumulation.  Biopython's _pwm.c sums in double and
//     casts once to float; summing fp32-rounded group entries in fp32 drifts past the 1e-4 contract
//     on sharp long motifs (|score| ~ 600: ulp 6e-5 per add).  Every motif therefore gets its own
//     power-of-two scale 2^e, as large as keeps any window sum inside int32; entries are
//     rint(sum * 2^e), integer adds are exact and order-independent, and the epilogue is one
//     int->float conversion (round to nearest, like the oracle's cast) and one exact multiply by
//     2^-e.  Error before the final rounding <= G/2 * 2^-e (1.9e-6 at |score| < 1000).  -inf entries
//     are a sentinel below every finite window sum (see tfbs_dense_quantise).
//   hits mode (scan_hits.cu): packed fixed-point filter + exact double re-scoring of candidates.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kScanThreads = 1024;                   // one persistent CTA per SM
constexpr int kPosPerThread = 4;
constexpr int kTile = kScanThreads * kPosPerThread;  // 4096 window starts per sub-tile

struct ScanParams {
    const uint32_t *packed;
    const uint32_t *mask;
    int64_t pitch;  // row pitch of the dense outputs (multiple of kTile)
    const int2 *lut;
    const float *dscale;    // [M] 2^-e of the motif's fixed-point entries
    const int32_t *dneg;    // [M] window sums below this mean -inf (INT32_MIN: motif has no -inf entry)
    const int32_t *lut_off;
    const int32_t *glen;
    int32_t M;
    float *out_fwd;
    float *out_rev;
};

// Dense mode.  Persistent CTAs; for each motif the k-mer table (k = 4 for L <= 24, else 3) is
// expanded into shared memory as T[g][x][lane & 15] (8 B = fwd, rev): an LDS.64 is served one
// half-warp at a time, so 16 private copies make every data-dependent load bank-conflict free (the
// v1 layout T[g][x] lost 60 % of its shared-memory wavefronts to conflicts,
// profiles/r01/ncu_full_v1_raw.csv) and leave room for 256-entry groups: L = 20 costs 5 loads per
// window instead of 7.  The table is 32 KB aligned so a lookup address is
// base | (kmer << 7) | ((lane & 15) << 3)  (one shift + one LOP3), the group offset is an
// immediate (K, G are template parameters) and the two strands are accumulated in int32 (ptxas fuses
// the adds of two groups into one IADD3).  The CTA streams over its sub-tiles of 8192 window starts: every thread scores 4
// consecutive starts from a 48-base register window and writes 16 B per strand (streaming stores,
// 512 contiguous bytes per warp).
__device__ __forceinline__ uint64_t lds64(uint32_t addr)
{
    uint64_t v;
    asm("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

template <int K, int G>
__device__ __forceinline__ void dense_tiles(const ScanParams &P, int m, int L, uint32_t lut_lane_addr)
{
    constexpr uint32_t kIndexMask = (uint32_t)((1 << (2 * K)) - 1) << 7;  // entry stride = 16 copies x 8 B
    constexpr int kGroupBytes = (1 << (2 * K)) * 128;
    const int tid = threadIdx.x;
    const int64_t n_tiles = P.pitch / kTile;
    const uint32_t lmask = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    float *row_f = P.out_fwd + (size_t)m * P.pitch;
    float *row_r = P.out_rev + (size_t)m * P.pitch;
    const float scale = __ldg(P.dscale + m);
    const int32_t neg_thr = __ldg(P.dneg + m);
    const float qnan = __uint_as_float(0x7FC00000u), ninf = __uint_as_float(0xFF800000u);

    // raw words are prefetched three tiles ahead (one round of the 32 resident warps is shorter
    // than an HBM round trip; ncu showed long-scoreboard stalls at smaller distances).  3 packed words (96 bases from a 16-base boundary) and 2 mask words cover the
    // 4 + 32 bases a thread can touch.
    struct Raw { uint32_t a0, a1, a2, m0, m1; };
    auto fetch = [&](int64_t tile) {
        Raw r;
        const int64_t q0 = tile * kTile + (int64_t)tid * kPosPerThread;
        const uint32_t *wp = P.packed + (q0 >> 4);
        r.a0 = __ldg(wp); r.a1 = __ldg(wp + 1); r.a2 = __ldg(wp + 2);
        const uint32_t *mp = P.mask + (q0 >> 5);
        r.m0 = __ldg(mp); r.m1 = __ldg(mp + 1);
        return r;
    };
    const int64_t stride = gridDim.x;
    // one tile: build the register window from raw words, score, store
    auto process = [&](int64_t tile, const Raw &cur) {
        const int64_t p0 = tile * kTile + (int64_t)tid * kPosPerThread;
        // 48-base register window starting at p0 (p0 is a multiple of 4)
        uint32_t w0, w1, w2, inv_lo, inv_hi;
        {
            const int sh = (int)(p0 & 15) * 2;
            w0 = __funnelshift_r(cur.a0, cur.a1, sh);
            w1 = __funnelshift_r(cur.a1, cur.a2, sh);
            w2 = cur.a2 >> sh;
            const int msh = (int)(p0 & 31);
            inv_lo = __funnelshift_r(cur.m0, cur.m1, msh);
            inv_hi = cur.m1 >> msh;
        }
        int32_t accf[kPosPerThread], accr[kPosPerThread];
#pragma unroll
        for (int g = 0; g < G; g++) {
#pragma unroll
            for (int p = 0; p < kPosPerThread; p++) {
                // bits [7, 7+2K) of t = k-mer code at base p0 + p + K*g
                const uint32_t t = (2 * p >= 7) ? (w0 >> (2 * p - 7)) : (w0 << (7 - 2 * p));
                const uint64_t v = lds64(((t & kIndexMask) | lut_lane_addr) + g * kGroupBytes);
                accf[p] = g == 0 ? (int32_t)(uint32_t)v : accf[p] + (int32_t)(uint32_t)v;
                accr[p] = g == 0 ? (int32_t)(v >> 32) : accr[p] + (int32_t)(v >> 32);
            }
            if (g + 1 < G) {
                w0 = __funnelshift_r(w0, w1, 2 * K);
                w1 = __funnelshift_r(w1, w2, 2 * K);
                w2 >>= 2 * K;
            }
        }
        float af[kPosPerThread], ar[kPosPerThread];
#pragma unroll
        for (int p = 0; p < kPosPerThread; p++) {
            const bool bad = (__funnelshift_r(inv_lo, inv_hi, p) & lmask) != 0u;
            const float f = accf[p] < neg_thr ? ninf : __int2float_rn(accf[p]) * scale;
            const float r = accr[p] < neg_thr ? ninf : __int2float_rn(accr[p]) * scale;
            af[p] = bad ? qnan : f;
            ar[p] = bad ? qnan : r;
        }
        // one fully contiguous 512 B store per warp and strand (a stride-32 B pattern writes half
        // sectors: twice the L1->XBAR sector traffic and store wavefronts, ncu r01)
        __stcs(reinterpret_cast<float4 *>(row_f + p0), make_float4(af[0], af[1], af[2], af[3]));
        __stcs(reinterpret_cast<float4 *>(row_r + p0), make_float4(ar[0], ar[1], ar[2], ar[3]));
    };
    // three register sets rotate without moves (a move would wait for the load it copies), so a
    // set is consumed two full iterations after its loads were issued
    const Raw zero = {0, 0, 0, 0, 0};
    int64_t tile = blockIdx.x;
    Raw r0 = tile < n_tiles ? fetch(tile) : zero;
    Raw r1 = tile + stride < n_tiles ? fetch(tile + stride) : zero;
    Raw r2 = tile + 2 * stride < n_tiles ? fetch(tile + 2 * stride) : zero;
    while (true) {
        if (tile >= n_tiles) break;
        process(tile, r0);
        if (tile + 3 * stride < n_tiles) r0 = fetch(tile + 3 * stride);
        tile += stride;
        if (tile >= n_tiles) break;
        process(tile, r1);
        if (tile + 3 * stride < n_tiles) r1 = fetch(tile + 3 * stride);
        tile += stride;
        if (tile >= n_tiles) break;
        process(tile, r2);
        if (tile + 3 * stride < n_tiles) r2 = fetch(tile + 3 * stride);
        tile += stride;
    }
}

__global__ void __launch_bounds__(kScanThreads, 1)
scan_dense_kernel(const ScanParams P)
{
    extern __shared__ uint8_t dense_smem_raw[];
    // 32 KB-align the table inside the dynamic shared window so that an entry address is
    // base | (kmer << 7) | ((lane & 15) << 3)
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(dense_smem_raw);
    const uint32_t pad = ((raw + 32767u) & ~32767u) - raw;
    int2 *s_lut = reinterpret_cast<int2 *>(dense_smem_raw + pad);  // [G][4^K][16]
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t lut_lane_addr = raw + pad + (uint32_t)(lane & 15) * 8u;

    for (int m = 0; m < P.M; m++) {
        const int32_t gl = __ldg(P.glen + m);
        const int K = gl >> 16, G = (gl >> 8) & 255, L = gl & 255;
        __syncthreads();  // everyone is done with the previous motif's table
        {
            const int2 *src = P.lut + __ldg(P.lut_off + m);
            const int total = G << (2 * K + 4);  // G * 4^K entries x 16 copies
            for (int i = tid; i < total; i += kScanThreads) s_lut[i] = __ldg(src + (i >> 4));
        }
        __syncthreads();
        switch ((K << 8) | G) {
#define TFBS_DCASE(k, g) \
    case ((k << 8) | g): dense_tiles<k, g>(P, m, L, lut_lane_addr); break;
            TFBS_DCASE(4, 1) TFBS_DCASE(4, 2) TFBS_DCASE(4, 3) TFBS_DCASE(4, 4) TFBS_DCASE(4, 5) TFBS_DCASE(4, 6)
            TFBS_DCASE(3, 9) TFBS_DCASE(3, 10) TFBS_DCASE(3, 11)
#undef TFBS_DCASE
            default: break;
        }
    }
}

void fill_params(ScanParams &P, const tfbs_seq *seq, const tfbs_pwms *pw, int64_t pitch)
{
    P.packed = seq->packed;
    P.mask = seq->mask;
    P.pitch = pitch;
    P.lut = pw->lut;
    P.dscale = pw->dense_scale;
    P.dneg = pw->dense_neg;
    P.lut_off = pw->lut_off;
    P.glen = pw->glen;
    P.M = pw->M;
    P.out_fwd = P.out_rev = nullptr;
}

}  // namespace

// Fixed-point k-mer tables of one motif for the dense kernel.  F / R: exact fwd / reverse-complement
// matrices [L][4].  k = 4 for L <= 24, else 3; entry (g, x) = sum of the <= k log-odds values of
// columns k*g.. selected by the k-mer x (first base in the low bits), in double.  The scale 2^e is the
// largest power of two (e <= 30) for which every window sum of rounded entries stays inside int32:
//   no -inf entry :  sum_g max_x |q| <= 2^31 - 1
//   with -inf     :  -inf groups become NEG = -(2 Rq + 1), Rq = sum_g max_x |q| over finite entries, so
//                    a window with a -inf group sums to <= NEG + Rq < -Rq <= any finite window sum and
//                    G * |NEG| + Rq must fit: the kernel reports -inf when the sum is < *neg_thr = -Rq.
// Appends G * 4^k int2 (fwd, rev) to `lut`.
void tfbs_dense_quantise(const double *F, const double *R, int L, int *K_out, int *G_out, int *e_out,
                         int32_t *neg_thr, std::vector<int2> &lut)
{
    const int K = L <= 24 ? 4 : 3, G = (L + K - 1) / K, E = 1 << (2 * K);
    std::vector<double> sf((size_t)G * E), sr((size_t)G * E);
    bool has_inf = false;
    double Rsum = 0.0;  // sum over groups of the largest finite |entry|
    for (int g = 0; g < G; g++) {
        double gmax = 0.0;
        for (int x = 0; x < E; x++) {
            double a = 0.0, b = 0.0;
            for (int j = 0; j < K && K * g + j < L; j++) {
                const int base = (x >> (2 * j)) & 3;
                a += F[(K * g + j) * 4 + base];
                b += R[(K * g + j) * 4 + base];
            }
            sf[(size_t)g * E + x] = a;
            sr[(size_t)g * E + x] = b;
            if (std::isinf(a)) has_inf = true; else gmax = std::max(gmax, std::fabs(a));
            if (std::isinf(b)) has_inf = true; else gmax = std::max(gmax, std::fabs(b));
        }
        Rsum += gmax;
    }
    // rounding adds at most 1/2 per group to a window sum
    const double room = has_inf ? (2147483647.0 - G) / (2.0 * G + 1.0) - G : 2147483647.0 - G;
    int e = 30;
    while (e > -60 && std::ldexp(Rsum, e) > room) e--;
    int64_t Rq = 0;
    std::vector<int64_t> qf((size_t)G * E), qr((size_t)G * E);
    for (int g = 0; g < G; g++) {
        int64_t gmax = 0;
        for (int x = 0; x < E; x++) {
            const double a = sf[(size_t)g * E + x], b = sr[(size_t)g * E + x];
            const int64_t ia = std::isinf(a) ? INT64_MIN : (int64_t)std::llrint(std::ldexp(a, e));
            const int64_t ib = std::isinf(b) ? INT64_MIN : (int64_t)std::llrint(std::ldexp(b, e));
            qf[(size_t)g * E + x] = ia;
            qr[(size_t)g * E + x] = ib;
            if (ia != INT64_MIN) gmax = std::max<int64_t>(gmax, ia < 0 ? -ia : ia);
            if (ib != INT64_MIN) gmax = std::max<int64_t>(gmax, ib < 0 ? -ib : ib);
        }
        Rq += gmax;
    }
    const int64_t NEG = -(2 * Rq + 1);
    for (size_t i = 0; i < qf.size(); i++)
        lut.push_back(make_int2((int)(qf[i] == INT64_MIN ? NEG : qf[i]), (int)(qr[i] == INT64_MIN ? NEG : qr[i])));
    *K_out = K;
    *G_out = G;
    *e_out = e;
    *neg_thr = has_inf ? (int32_t)(-Rq) : INT32_MIN;
}

// Hits mode, library-level host state: motifs sorted by descending length, and the capacity of the device
// arrays.  The cut into blocks (32 motifs with 16-bit filter fields or 64 with 8-bit fields) depends on the
// thresholds and is made with the filter tables (plan_hits_tables in scan_hits.cu); the arrays are sized for
// the largest plan, 32 motifs per block throughout.  A per-threshold plan leaves some short motifs to the
// direct path, so block i may start at a motif SHORTER than lens_desc[32 i]; the table of a shorter motif
// can be the larger one — 24 columns take 6 x 32 KB, 25 columns 5 x 32 + 2 x 8 KB — hence the running
// maximum over all lengths up to it.
void tfbs_hits_order_and_capacity(tfbs_pwms *p, const int32_t *lens)
{
    const int M = p->M;
    p->order.resize(M);
    for (int m = 0; m < M; m++) p->order[m] = m;
    std::stable_sort(p->order.begin(), p->order.end(), [&](int a, int b) { return lens[a] > lens[b]; });
    p->lens_desc.resize(M);
    for (int i = 0; i < M; i++) p->lens_desc[i] = lens[p->order[i]];
    p->n_blocks = 0;
    p->n_blocks_max = (M + 31) / 32;
    int64_t words = 0, words_upto[TFBS_MAX_MOTIF_LEN + 1] = {0};
    for (int L = 1; L <= TFBS_MAX_MOTIF_LEN; L++) {
        int A = 0, B = 1;
        tfbs_hits_group_plan(L, &A, &B);
        words_upto[L] = std::max(words_upto[L - 1], (int64_t)A * 256 * 32 + (int64_t)B * 64 * 32);
    }
    for (int i = 0; i < M; i += 32) words += words_upto[p->lens_desc[i]];
    p->hits_lut_bytes = (size_t)words_upto[p->lens_desc[0]] * 4;
    p->hq_lut_words = words;
}

extern "C" {

int tfbs_debug_dense_tables(const double *logodds, int32_t L, int32_t *K, int32_t *G, int32_t *scale_exp,
                            int32_t *neg_thr, int32_t *entries, int64_t cap)
{
    TFBS_REQUIRE(logodds && K && G && scale_exp && neg_thr && L >= 1 && L <= TFBS_MAX_MOTIF_LEN,
                 "tfbs_debug_dense_tables: bad argument");
    double F[TFBS_MAX_MOTIF_LEN * 4], R[TFBS_MAX_MOTIF_LEN * 4];
    tfbs_fill_exact(logodds, L, 0, L, F, R);
    std::vector<int2> lut;
    int k = 0, g = 0, e = 0;
    tfbs_dense_quantise(F, R, L, &k, &g, &e, neg_thr, lut);
    *K = k; *G = g; *scale_exp = e;
    if (entries) {
        TFBS_REQUIRE(cap >= (int64_t)lut.size() * 2, "tfbs_debug_dense_tables: entries holds %lld < %lld",
                     (long long)cap, (long long)lut.size() * 2);
        for (size_t i = 0; i < lut.size(); i++) { entries[2 * i] = lut[i].x; entries[2 * i + 1] = lut[i].y; }
    }
    return TFBS_OK;
}

int tfbs_pwm_upload(tfbs_ctx *ctx, const double *logodds, const int32_t *lens, int32_t M,
                    int32_t Lmax, tfbs_pwms **out)
{
    TFBS_REQUIRE(ctx && logodds && lens && out && M > 0 && Lmax > 0, "tfbs_pwm_upload: bad argument");
    TFBS_REQUIRE(Lmax <= TFBS_MAX_MOTIF_LEN, "tfbs_pwm_upload: Lmax %d > %d", Lmax, TFBS_MAX_MOTIF_LEN);
    for (int m = 0; m < M; m++) {
        TFBS_REQUIRE(lens[m] >= 1 && lens[m] <= Lmax, "tfbs_pwm_upload: motif %d has length %d", m, lens[m]);
        for (int j = 0; j < lens[m]; j++)
            for (int b = 0; b < 4; b++) {
                const double v = logodds[((size_t)m * Lmax + j) * 4 + b];
                TFBS_REQUIRE(!(std::isnan(v) || (std::isinf(v) && v > 0)),
                             "tfbs_pwm_upload: motif %d column %d has a NaN/+inf entry", m, j);
            }
    }
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));

    tfbs_pwms *p = new tfbs_pwms();
    p->ctx = ctx;
    p->M = M;
    p->Lmax = Lmax;
    p->lens.assign(lens, lens + M);
    p->logodds.assign(logodds, logodds + (size_t)M * Lmax * 4);

    std::vector<double> exact((size_t)M * 2 * TFBS_MAX_MOTIF_LEN * 4, 0.0);
    std::vector<int32_t> lut_off(M), glen(M), dneg(M);
    std::vector<float> dscale(M);
    std::vector<int2> lut;
    int Gmax = 0;
    size_t dense_lut_bytes = 0;  // largest per-motif table once expanded to 16 lane copies
    for (int m = 0; m < M; m++) {
        const int L = lens[m], G = (L + 2) / 3;
        Gmax = std::max(Gmax, G);
        double *F = &exact[((size_t)m * 2 + 0) * TFBS_MAX_MOTIF_LEN * 4];
        double *R = &exact[((size_t)m * 2 + 1) * TFBS_MAX_MOTIF_LEN * 4];
        tfbs_fill_exact(logodds, Lmax, m, L, F, R);
        // dense mode: fixed-point k-mer tables, k = 4 (256 entries per group) for L <= 24, else k = 3
        int K = 0, Gd = 0, e = 0;
        lut_off[m] = (int32_t)lut.size();
        tfbs_dense_quantise(F, R, L, &K, &Gd, &e, &dneg[m], lut);
        dscale[m] = std::ldexp(1.0f, -e);
        glen[m] = (K << 16) | (Gd << 8) | L;
        dense_lut_bytes = std::max(dense_lut_bytes, (size_t)Gd * (1 << (2 * K)) * 16 * sizeof(int2));
    }
    // pair tables + rounding guard of the verify pass (scan_hits.cu: verify_candidates_kernel)
    std::vector<double> exact2((size_t)M * 2 * 16 * 16, 0.0), guard(M, 0.0);
    for (int m = 0; m < M; m++) {
        const int L = lens[m];
        double A = 0.0;
        for (int s2 = 0; s2 < 2; s2++) {
            const double *mat = &exact[((size_t)m * 2 + s2) * TFBS_MAX_MOTIF_LEN * 4];
            double *dst = &exact2[((size_t)m * 2 + s2) * 256];
            for (int j = 0; 2 * j < L; j++)
                for (int x = 0; x < 16; x++) {
                    const double a = mat[(2 * j) * 4 + (x & 3)];
                    const double b = 2 * j + 1 < L ? mat[(2 * j + 1) * 4 + (x >> 2)] : 0.0;
                    dst[j * 16 + x] = a + b;
                }
            if (s2 == 0)
                for (int j = 0; j < L; j++) {
                    double mx = 0.0;
                    for (int b = 0; b < 4; b++)
                        if (std::isfinite(mat[j * 4 + b])) mx = std::max(mx, std::fabs(mat[j * 4 + b]));
                    A += mx;
                }
        }
        guard[m] = std::ldexp(A, -46);
    }
    p->dense_lut_bytes = dense_lut_bytes;
    p->Gmax = Gmax;
    p->lut_entries = (int64_t)lut.size();

    // hits mode: length-sorted motif order and the capacity of the per-threshold device arrays
    tfbs_hits_order_and_capacity(p, lens);
    p->exact_host = exact;

    cudaError_t e = cudaMalloc(&p->exact, exact.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&p->exact2, exact2.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&p->guard_eps, guard.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(p->exact2, exact2.data(), exact2.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->guard_eps, guard.data(), guard.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->lut, lut.size() * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc(&p->dense_scale, M * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&p->dense_neg, M * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->lut_off, M * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->glen, M * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->hq_lut_dev, (size_t)p->hq_lut_words * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->hq_kinit_dev, (size_t)p->n_blocks_max * 32 * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->blk_AB_dev, (size_t)p->n_blocks_max * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->blk_lut_off_dev, (size_t)p->n_blocks_max * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->blk_motif_dev, (size_t)p->n_blocks_max * TFBS_BLOCK_SLOTS * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->lens_dev, (size_t)M * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->lens_dev, lens, (size_t)M * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->exact, exact.data(), exact.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->lut, lut.data(), lut.size() * sizeof(int2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->dense_scale, dscale.data(), M * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->dense_neg, dneg.data(), M * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->lut_off, lut_off.data(), M * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->glen, glen.data(), M * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        tfbs_set_error("tfbs_pwm_upload: CUDA error: %s", cudaGetErrorString(e));
        tfbs_pwms_destroy(p);
        return e == cudaErrorMemoryAllocation ? TFBS_ERR_NOMEM : TFBS_ERR_CUDA;
    }
    *out = p;
    return TFBS_OK;
}

int tfbs_pwms_destroy(tfbs_pwms *p)
{
    if (!p) return TFBS_OK;
    if (p->ctx) cudaSetDevice(p->ctx->device);
    if (p->exact) cudaFree(p->exact);
    if (p->exact2) cudaFree(p->exact2);
    if (p->guard_eps) cudaFree(p->guard_eps);
    if (p->lut) cudaFree(p->lut);
    if (p->dense_scale) cudaFree(p->dense_scale);
    if (p->dense_neg) cudaFree(p->dense_neg);
    if (p->lut_off) cudaFree(p->lut_off);
    if (p->glen) cudaFree(p->glen);
    if (p->hq_lut_dev) cudaFree(p->hq_lut_dev);
    if (p->direct_bitmap_dev) cudaFree(p->direct_bitmap_dev);
    if (p->direct_hash_dev) cudaFree(p->direct_hash_dev);
    if (p->hq_kinit_dev) cudaFree(p->hq_kinit_dev);
    if (p->blk_AB_dev) cudaFree(p->blk_AB_dev);
    if (p->blk_lut_off_dev) cudaFree(p->blk_lut_off_dev);
    if (p->blk_motif_dev) cudaFree(p->blk_motif_dev);
    if (p->lens_dev) cudaFree(p->lens_dev);
    delete p;
    return TFBS_OK;
}

int tfbs_scan_dense(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_pwms *pwms, float *fwd_host,
                    float *rev_host)
{
    TFBS_REQUIRE(ctx && seq && pwms, "tfbs_scan_dense: null argument");
    TFBS_REQUIRE(seq->ctx == ctx && pwms->ctx == ctx, "tfbs_scan_dense: handle from another context");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = seq->n;
    if (n == 0) return TFBS_OK;
    const int64_t pitch = ((n + kTile - 1) / kTile) * kTile;
    const size_t bytes = (size_t)pwms->M * pitch * sizeof(float);
    int rc = tfbs_scratch_reserve(ctx->dense_fwd, bytes);
    if (rc) return rc;
    rc = tfbs_scratch_reserve(ctx->dense_rev, bytes);
    if (rc) return rc;
    ScanParams P;
    fill_params(P, seq, pwms, pitch);
    P.out_fwd = (float *)ctx->dense_fwd.ptr;
    P.out_rev = (float *)ctx->dense_rev.ptr;
    const size_t smem = pwms->dense_lut_bytes + 32768;  // + alignment slack
    TFBS_CHECK_CUDA(cudaFuncSetAttribute(scan_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = std::min<int64_t>(ctx->sm_count, pitch / kTile);
    tfbs_begin(ctx);
    scan_dense_kernel<<<(unsigned)grid, kScanThreads, smem, ctx->stream>>>(P);
    tfbs_launched(ctx);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    if (fwd_host)
        TFBS_CHECK_CUDA(cudaMemcpy2DAsync(fwd_host, (size_t)n * 4, ctx->dense_fwd.ptr, (size_t)pitch * 4,
                                          (size_t)n * 4, pwms->M, cudaMemcpyDeviceToHost, ctx->stream));
    if (rev_host)
        TFBS_CHECK_CUDA(cudaMemcpy2DAsync(rev_host, (size_t)n * 4, ctx->dense_rev.ptr, (size_t)pitch * 4,
                                          (size_t)n * 4, pwms->M, cudaMemcpyDeviceToHost, ctx->stream));
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return TFBS_OK;
}

}  // extern "C"
