// shape.cu — (3) DNA-shape pentamer-table lookup.
//
// Replaces DNAshapeR getShape()/getDNAShape() (get_DNA_shapes.R:23-25,
// create_bkg_seqs.py:365-366), the 71.5 M-string parse of its text output
// (build_features.py:193-198) and the per-site window slicing of utils.py:1061-1084
// (_get_DNA_shape) and build_features.py:215-241 (background trims).
//
// The pentamer tables (1024 entries per value plane) live in shared memory, indexed by the
// little-endian pentamer index that falls straight out of the 2-bit packed words
// (x = sum code[c-2+k] << 2k); the host re-indexes the caller's big-endian tables once.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

constexpr int kShapeThreads = 256;
constexpr int kShapePosPerThread = 4;
constexpr int kShapeTile = 8192;  // positions per CTA iteration (8 sub-tiles of 1024)
constexpr int kMaxTracks = 8;

struct TrackParams {
    const uint32_t *packed;
    const uint32_t *mask;
    int64_t n_words;  // allocation sizes, for the guards at the buffer start
    int64_t n_mask;
    int64_t pitch;
    const float *tables;  // [T][2][1024], little-endian pentamer index
    int32_t T;
    uint32_t step_mask;  // bit t set: track t is per-step
    float *out;  // [T][pitch]
};

// Streaming kernel.  Persistent CTAs keep the tables in shared memory; a lane produces 4
// consecutive positions of every track, so each track costs one fully contiguous STG.128 per warp
// (a stride-32 B pattern doubled the store wavefronts in the l1tex data pipe, which is the unit
// this kernel saturates: profiles/r01/ncu_full_dense_shape_encode_v2_100Mb_raw.csv).  The
// pentamer just past a lane's 4 positions (needed by the last step value) comes from the next
// lane by shuffle; only lane 31 looks it up itself.
__global__ void __launch_bounds__(kShapeThreads)
shape_tracks_kernel(const TrackParams P)
{
    extern __shared__ float s_tab[];  // [T][2][1024]
    const int tid = threadIdx.x, lane = tid & 31;
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.tables);
        float4 *dst = reinterpret_cast<float4 *>(s_tab);
        for (int i = tid; i < P.T * 512; i += kShapeThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const float qnan = __int_as_float(0x7FC00000);
    const int64_t n_sub = P.pitch / (kShapeThreads * kShapePosPerThread);
    // raw words are prefetched three sub-tiles ahead (hides the HBM round trip)
    struct Raw { uint32_t a0, a1, m0, m1; };
    auto fetch = [&](int64_t sub) {
        Raw r;
        const int64_t q0 = sub * (kShapeThreads * kShapePosPerThread) + tid * kShapePosPerThread - 4;
        // q0 >= -4: the encoded buffers carry a front pad, so index -1 is readable (invalid bases)
        const uint32_t *wp = P.packed + (q0 >> 4);
        const uint32_t *mp = P.mask + (q0 >> 5);
        r.a0 = __ldg(wp);
        r.a1 = __ldg(wp + 1);
        r.m0 = __ldg(mp);
        r.m1 = __ldg(mp + 1);
        return r;
    };
    const int64_t stride = gridDim.x;
    auto process = [&](int64_t sub, const Raw &cur) {
        const int64_t p0 = sub * (kShapeThreads * kShapePosPerThread) + tid * kShapePosPerThread;
        // 16-base register window starting at base b0 = p0 - 4 (a multiple of 4, may be -4):
        // pentamer centred at p0 + d starts at window base d + 2
        const int64_t b0 = p0 - 4;
        const uint32_t wlo = __funnelshift_r(cur.a0, cur.a1, (int)(b0 & 15) * 2);
        const uint32_t inv = __funnelshift_r(cur.m0, cur.m1, (int)(b0 & 31));
        uint32_t pent[kShapePosPerThread + 1];
        bool ok[kShapePosPerThread + 1];
#pragma unroll
        for (int d = 0; d <= kShapePosPerThread; d++) {
            pent[d] = (wlo >> (2 * (d + 2))) & 1023u;
            ok[d] = ((inv >> (d + 2)) & 31u) == 0u;
        }
        for (int t = 0; t < P.T; t++) {
            const float *y1 = s_tab + t * 2048;
            const float *y2 = y1 + 1024;
            float v[kShapePosPerThread];
            if (((P.step_mask >> t) & 1u) == 0u) {
#pragma unroll
                for (int d = 0; d < kShapePosPerThread; d++) v[d] = ok[d] ? y1[pent[d]] : qnan;
            } else {
                // step k between bases k and k+1: mean of Y2(pentamer@k) and Y1(pentamer@k+1).
                // (a + b) * 0.5f in fp32 equals the oracle's (float)(((double)a + (double)b) / 2).
                float a1[kShapePosPerThread + 1];
#pragma unroll
                for (int d = 0; d < kShapePosPerThread; d++) a1[d] = y1[pent[d]];
                a1[kShapePosPerThread] = __shfl_down_sync(0xFFFFFFFFu, a1[0], 1);
                if (lane == 31) a1[kShapePosPerThread] = y1[pent[kShapePosPerThread]];
#pragma unroll
                for (int d = 0; d < kShapePosPerThread; d++)
                    v[d] = (ok[d] && ok[d + 1]) ? __fmul_rn(__fadd_rn(y2[pent[d]], a1[d + 1]), 0.5f) : qnan;
            }
            __stcs(reinterpret_cast<float4 *>(P.out + (size_t)t * P.pitch + p0), make_float4(v[0], v[1], v[2], v[3]));
        }
    };
    // three register sets rotate without moves (a move would wait for the load it copies)
    const Raw none = {0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu};
    int64_t sub = blockIdx.x;
    Raw r0 = sub < n_sub ? fetch(sub) : none;
    Raw r1 = sub + stride < n_sub ? fetch(sub + stride) : none;
    Raw r2 = sub + 2 * stride < n_sub ? fetch(sub + 2 * stride) : none;
    while (true) {
        if (sub >= n_sub) break;
        process(sub, r0);
        if (sub + 3 * stride < n_sub) r0 = fetch(sub + 3 * stride);
        sub += stride;
        if (sub >= n_sub) break;
        process(sub, r1);
        if (sub + 3 * stride < n_sub) r1 = fetch(sub + 3 * stride);
        sub += stride;
        if (sub >= n_sub) break;
        process(sub, r2);
        if (sub + 3 * stride < n_sub) r2 = fetch(sub + 3 * stride);
        sub += stride;
    }
}

// ---- packed centi-unit variant ---------------------------------------------------------------------
// DNA-shape tables hold 2-decimal values (DNAshapeR prints %.2f).  When every entry v satisfies
// v == (float)((double)c * 0.01) for an integer |c| < 32768 (checked at upload), two values share one
// 32-bit word: (track A | track B) for per-nucleotide tracks, (Y1 | Y2) for a per-step track.  That
// halves the data-dependent LDS.32 per position — the bank conflicts of those loads are what
// saturates the l1tex data pipe in the fp32-table kernel (71 % of its shared-memory wavefronts,
// profiles/r01/ncu_full_final_100Mb_raw.csv).  The reconstruction (centi_from_biased) is verified
// against every table entry at upload.
struct Packed16Params {
    const uint32_t *packed;
    const uint32_t *mask;
    int64_t n_words, n_mask, pitch;
    const uint32_t *words;  // [NW][1024] little-endian pentamer index
    int32_t NW;
    int32_t kind[kMaxTracks];   // 0: two per-nucleotide tracks, 1: one per-step track (Y1 | Y2 << 16)
    int64_t off_a[kMaxTracks];  // row * pitch of the low half (or of the step track)
    int64_t off_b[kMaxTracks];  // row * pitch of the high half, -1 if none
    float *out;
};

// c / 100 correctly rounded to float with 4 full-rate ops (int->float and float<->double
// conversions run on the quarter-rate pipe).  The table stores the biased value h = c + 32768 in 16
// bits; OR-ing it under the exponent of 2^23 gives the float 2^23 + h exactly, one subtraction gives
// c, and c * 0.01 is formed as fma(c, r_hi, c * r_lo) with 0.01 = r_hi + r_lo.  The host runs the same
// sequence for every table entry at upload and only enables this kernel if it reproduces the table
// bit for bit.
__host__ __device__ __forceinline__ float centi_from_biased(uint32_t h)  // h in [0, 65535]
{
    const float r_hi = 0.01f, r_lo = 2.2351741811588166e-10f;  // 0.01 - (double)0.01f
#ifdef __CUDA_ARCH__
    const float cf = __uint_as_float(0x4B000000u | h) - 8421376.0f;  // (2^23 + h) - (2^23 + 32768)
    return __fmaf_rn(cf, r_hi, __fmul_rn(cf, r_lo));
#else
    const float cf = (float)((int)h - 32768);
    return fmaf(cf, r_hi, cf * r_lo);
#endif
}

__global__ void __launch_bounds__(kShapeThreads)
shape_tracks_c16_kernel(const Packed16Params P)
{
    extern __shared__ uint32_t s_w[];  // [NW][1024]
    const int tid = threadIdx.x, lane = tid & 31;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.words);
        uint4 *dst = reinterpret_cast<uint4 *>(s_w);
        for (int i = tid; i < P.NW * 256; i += kShapeThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const float qnan = __int_as_float(0x7FC00000);
    const int64_t n_sub = P.pitch / (kShapeThreads * kShapePosPerThread);
    // raw words are prefetched three sub-tiles ahead (hides the HBM round trip)
    struct Raw { uint32_t a0, a1, m0, m1; };
    auto fetch = [&](int64_t sub) {
        Raw r;
        const int64_t q0 = sub * (kShapeThreads * kShapePosPerThread) + tid * kShapePosPerThread - 4;
        // q0 >= -4: the encoded buffers carry a front pad, so index -1 is readable (invalid bases)
        const uint32_t *wp = P.packed + (q0 >> 4);
        const uint32_t *mp = P.mask + (q0 >> 5);
        r.a0 = __ldg(wp);
        r.a1 = __ldg(wp + 1);
        r.m0 = __ldg(mp);
        r.m1 = __ldg(mp + 1);
        return r;
    };
    const int64_t stride = gridDim.x;
    auto process = [&](int64_t sub, const Raw &cur) {
        const int64_t p0 = sub * (kShapeThreads * kShapePosPerThread) + tid * kShapePosPerThread;
        const int64_t b0 = p0 - 4;
        const uint32_t wlo = __funnelshift_r(cur.a0, cur.a1, (int)(b0 & 15) * 2);
        const uint32_t inv = __funnelshift_r(cur.m0, cur.m1, (int)(b0 & 31));
        uint32_t pent[kShapePosPerThread + 1];
        bool ok[kShapePosPerThread + 1];
#pragma unroll
        for (int d = 0; d <= kShapePosPerThread; d++) {
            pent[d] = (wlo >> (2 * (d + 2))) & 1023u;
            ok[d] = ((inv >> (d + 2)) & 31u) == 0u;
        }
        for (int w = 0; w < P.NW; w++) {
            const uint32_t *tab = s_w + w * 1024;
            uint32_t u[kShapePosPerThread + 1];
#pragma unroll
            for (int d = 0; d < kShapePosPerThread; d++) u[d] = tab[pent[d]];
            if (P.kind[w] == 0) {
                float va[kShapePosPerThread], vb[kShapePosPerThread];
#pragma unroll
                for (int d = 0; d < kShapePosPerThread; d++) {
                    va[d] = ok[d] ? centi_from_biased(u[d] & 0xFFFFu) : qnan;
                    vb[d] = ok[d] ? centi_from_biased(u[d] >> 16) : qnan;
                }
                __stcs(reinterpret_cast<float4 *>(P.out + P.off_a[w] + p0), make_float4(va[0], va[1], va[2], va[3]));
                if (P.off_b[w] >= 0)
                    __stcs(reinterpret_cast<float4 *>(P.out + P.off_b[w] + p0), make_float4(vb[0], vb[1], vb[2], vb[3]));
            } else {
                // the pentamer just past this lane's 4 positions belongs to the next lane
                u[kShapePosPerThread] = __shfl_down_sync(0xFFFFFFFFu, u[0], 1);
                if (lane == 31) u[kShapePosPerThread] = tab[pent[kShapePosPerThread]];
                float v[kShapePosPerThread];
#pragma unroll
                for (int d = 0; d < kShapePosPerThread; d++) {
                    const float y2 = centi_from_biased(u[d] >> 16);
                    const float y1 = centi_from_biased(u[d + 1] & 0xFFFFu);
                    v[d] = (ok[d] && ok[d + 1]) ? __fmul_rn(__fadd_rn(y2, y1), 0.5f) : qnan;
                }
                __stcs(reinterpret_cast<float4 *>(P.out + P.off_a[w] + p0), make_float4(v[0], v[1], v[2], v[3]));
            }
        }
    };
    // three register sets rotate without moves (a move would wait for the load it copies)
    const Raw none = {0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu};
    int64_t sub = blockIdx.x;
    Raw r0 = sub < n_sub ? fetch(sub) : none;
    Raw r1 = sub + stride < n_sub ? fetch(sub + stride) : none;
    Raw r2 = sub + 2 * stride < n_sub ? fetch(sub + 2 * stride) : none;
    while (true) {
        if (sub >= n_sub) break;
        process(sub, r0);
        if (sub + 3 * stride < n_sub) r0 = fetch(sub + 3 * stride);
        sub += stride;
        if (sub >= n_sub) break;
        process(sub, r1);
        if (sub + 3 * stride < n_sub) r1 = fetch(sub + 3 * stride);
        sub += stride;
        if (sub >= n_sub) break;
        process(sub, r2);
        if (sub + 3 * stride < n_sub) r2 = fetch(sub + 3 * stride);
        sub += stride;
    }
}

// Value of one track at index idx of a record [rs, rs+rl) — the DNAshapeR rule including the
// single-pentamer step values next to the record ends.  Returns NaN for NA.
__device__ __forceinline__ int pentamer_at(const uint32_t *__restrict__ packed,
                                           const uint32_t *__restrict__ mask, int64_t rs, int64_t rl,
                                           int64_t c, uint32_t *idx)
{
    if (c < 2 || c + 2 >= rl) return -2;  // does not exist
    uint32_t x = 0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int64_t i = rs + c - 2 + k;
        bad |= ((__ldg(mask + (i >> 5)) >> (i & 31)) & 1u) != 0u;
        x |= ((__ldg(packed + (i >> 4)) >> (2 * (i & 15))) & 3u) << (2 * k);
    }
    *idx = x;
    return bad ? -1 : 0;
}

__device__ float shape_value(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask,
                             const float *__restrict__ tab, int kind, int64_t rs, int64_t rl,
                             int64_t idx)
{
    const float qnan = __int_as_float(0x7FC00000);
    const float *y1 = tab, *y2 = tab + 1024;
    uint32_t xa = 0, xb = 0;
    if (kind == TFBS_SHAPE_PER_NUCLEOTIDE) {
        const int ra = pentamer_at(packed, mask, rs, rl, idx, &xa);
        return ra == 0 ? __ldg(y1 + xa) : qnan;
    }
    const int ra = pentamer_at(packed, mask, rs, rl, idx, &xa);      // contributes Y2
    const int rb = pentamer_at(packed, mask, rs, rl, idx + 1, &xb);  // contributes Y1
    if (ra == -1 || rb == -1 || (ra == -2 && rb == -2)) return qnan;
    if (ra == -2) return __ldg(y1 + xb);
    if (rb == -2) return __ldg(y2 + xa);
    return __fmul_rn(__fadd_rn(__ldg(y2 + xa), __ldg(y1 + xb)), 0.5f);
}

// Patches the two single-pentamer step values of every record (k = 1 and k = len-3) that the
// streaming kernel, which only sees the invalid mask, had to leave as NA.
__global__ void __launch_bounds__(256)
shape_edge_fix_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask,
                      const float *__restrict__ tables, int32_t T, const int32_t *__restrict__ kinds,
                      const int64_t *__restrict__ rec_start, const int64_t *__restrict__ rec_len,
                      int64_t S, int64_t pitch, float *__restrict__ out)
{
    const int64_t total = S * 2 * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const int side = (int)((i / T) & 1);
        const int64_t rec = i / (2 * T);
        if (__ldg(kinds + t) != TFBS_SHAPE_PER_STEP) continue;
        const int64_t rs = __ldg(rec_start + rec), rl = __ldg(rec_len + rec);
        if (rl < 5) continue;
        const int64_t k = side == 0 ? 1 : rl - 3;
        out[(size_t)t * pitch + rs + k] =
            shape_value(packed, mask, tables + (size_t)t * 2048, TFBS_SHAPE_PER_STEP, rs, rl, k);
    }
}

// out[s][t][w]: the reference's per-site slices (see tfbs_cuda.h for the modes).
__global__ void __launch_bounds__(256)
shape_sites_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask,
                   const float *__restrict__ tables, int32_t T, const int32_t *__restrict__ kinds,
                   const int64_t *__restrict__ rec_start, const int64_t *__restrict__ rec_len,
                   const int64_t *__restrict__ start, const int64_t *__restrict__ end,
                   const int8_t *__restrict__ strand, int64_t S, int32_t W, int32_t mode,
                   int32_t trailing_empty, float *__restrict__ out)
{
    const float qnan = __int_as_float(0x7FC00000);
    const int64_t total = S * T * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int t = (int)((i / W) % T);
        const int64_t s = i / ((int64_t)W * T);
        const int kind = __ldg(kinds + t);
        const int64_t rs = __ldg(rec_start + s), rl = __ldg(rec_len + s);
        const int64_t n_tok = kind == TFBS_SHAPE_PER_NUCLEOTIDE ? rl : rl - 1;  // real tokens
        int64_t idx;
        bool present;
        if (mode == 0) {
            const int64_t a = __ldg(start + s), b = __ldg(end + s);
            const int64_t list_len = n_tok + (trailing_empty ? 1 : 0);
            if (__ldg(strand + s) >= 0) {
                // Python slice [a : b+1]
                idx = a + w;
                present = a >= 0 && idx <= b && idx < list_len;
            } else {
                // Python slice [b : a-1 : -1]; a == 0 makes the stop -1 == "last element", i.e. an
                // empty slice; a start beyond the list is clamped to the last element
                const int64_t first = b < list_len - 1 ? b : list_len - 1;
                idx = first - w;
                present = a >= 1 && idx >= a && list_len > 0;
            }
        } else {
            // background trims: per-nucleotide [2:-2] of rl tokens, per-step [1:-1] of rl-1 tokens
            const int64_t lo = kind == TFBS_SHAPE_PER_NUCLEOTIDE ? 2 : 1;
            const int64_t hi = kind == TFBS_SHAPE_PER_NUCLEOTIDE ? rl - 2 : rl - 2;  // exclusive
            idx = lo + w;
            present = idx < hi;
        }
        float v = qnan;
        if (present) {
            if (idx >= n_tok) v = 0.0f;  // the trailing '' of the flattened file -> 0.0
            else {
                v = shape_value(packed, mask, tables + (size_t)t * 2048, kind, rs, rl, idx);
                if (v != v) v = 0.0f;    // 'NA' -> ValueError -> 0.0 (utils.py:1076-1077)
            }
        }
        out[i] = v;
    }
}


// Python-slice gather over precomputed per-record token values (legacy input: the reference's
// {shape: {chrom: [tokens]}} dict, already parsed to floats with NA/'' -> 0.0 on the host):
// out[s][w] = values[rec_off[s] + idx] for idx in the reference's slice, NaN where the slice is
// shorter than W (utils.py:1061-1084).
__global__ void __launch_bounds__(256)
track_sites_kernel(const float *__restrict__ values, const int64_t *__restrict__ rec_off,
                   const int64_t *__restrict__ rec_ntok, const int64_t *__restrict__ start,
                   const int64_t *__restrict__ end, const int8_t *__restrict__ strand, int64_t S,
                   int32_t W, float *__restrict__ out)
{
    const float qnan = __int_as_float(0x7FC00000);
    const int64_t total = S * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int64_t s = i / W;
        const int64_t a = __ldg(start + s), b = __ldg(end + s), list_len = __ldg(rec_ntok + s);
        int64_t idx;
        bool present;
        if (__ldg(strand + s) >= 0) {
            idx = a + w;
            present = a >= 0 && idx <= b && idx < list_len;
        } else {
            const int64_t first = b < list_len - 1 ? b : list_len - 1;
            idx = first - w;
            present = a >= 1 && idx >= a && list_len > 0;
        }
        out[i] = present ? __ldg(values + __ldg(rec_off + s) + idx) : qnan;
    }
}

}  // namespace

extern "C" {

int tfbs_shapes_upload(tfbs_ctx *ctx, const float *tables, const int32_t *kinds, int32_t T,
                       tfbs_shapes **out)
{
    TFBS_REQUIRE(ctx && tables && kinds && out && T > 0 && T <= kMaxTracks,
                 "tfbs_shapes_upload: bad argument (T must be 1..%d)", kMaxTracks);
    for (int t = 0; t < T; t++)
        TFBS_REQUIRE(kinds[t] == TFBS_SHAPE_PER_NUCLEOTIDE || kinds[t] == TFBS_SHAPE_PER_STEP,
                     "tfbs_shapes_upload: kinds[%d] = %d", t, kinds[t]);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    tfbs_shapes *s = new tfbs_shapes();
    s->ctx = ctx;
    s->T = T;
    s->kinds.assign(kinds, kinds + T);
    // re-index: little-endian index x (first base in the low bits) <- big-endian index
    std::vector<float> le((size_t)T * 2048);
    for (int t = 0; t < T; t++)
        for (int plane = 0; plane < 2; plane++)
            for (int x = 0; x < 1024; x++) {
                int be = 0;
                for (int k = 0; k < 5; k++) be = be * 4 + ((x >> (2 * k)) & 3);
                le[((size_t)t * 2 + plane) * 1024 + x] = tables[((size_t)t * 2 + plane) * 1024 + be];
            }
    // packed centi-unit planes (see shape_tracks_c16_kernel)
    {
        bool centi = true;
        std::vector<int32_t> cv(le.size());
        for (size_t i = 0; i < le.size() && centi; i++) {
            const long c = lrintf(le[i] * 100.0f);
            cv[i] = (int32_t)c;
            if (c < -32767 || c > 32767 || centi_from_biased((uint32_t)(c + 32768)) != le[i]) centi = false;
        }
        if (centi) {
            std::vector<int> pern, step;
            for (int t = 0; t < T; t++) (kinds[t] == TFBS_SHAPE_PER_NUCLEOTIDE ? pern : step).push_back(t);
            std::vector<uint32_t> words;
            auto plane = [&](int t, int pl, int x) { return (uint32_t)(cv[((size_t)t * 2 + pl) * 1024 + x] + 32768) & 0xFFFFu; };
            for (size_t i = 0; i < pern.size(); i += 2) {
                const int a = pern[i], b = i + 1 < pern.size() ? pern[i + 1] : -1;
                s->c16_kind.push_back(0);
                s->c16_row_a.push_back(a);
                s->c16_row_b.push_back(b);
                for (int x = 0; x < 1024; x++) words.push_back(plane(a, 0, x) | (b >= 0 ? plane(b, 0, x) << 16 : 0u));
            }
            for (int t : step) {
                s->c16_kind.push_back(1);
                s->c16_row_a.push_back(t);
                s->c16_row_b.push_back(-1);
                for (int x = 0; x < 1024; x++) words.push_back(plane(t, 0, x) | (plane(t, 1, x) << 16));
            }
            cudaError_t e2 = cudaMalloc(&s->c16_words, words.size() * 4);
            if (e2 == cudaSuccess) e2 = cudaMemcpy(s->c16_words, words.data(), words.size() * 4, cudaMemcpyHostToDevice);
            if (e2 != cudaSuccess) {
                cudaGetLastError();
                if (s->c16_words) cudaFree(s->c16_words);
                s->c16_words = nullptr;
            }
        }
    }
    cudaError_t e = cudaMalloc(&s->tables, le.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->kinds_dev, T * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(s->tables, le.data(), le.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->kinds_dev, kinds, T * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        tfbs_set_error("tfbs_shapes_upload: CUDA error: %s", cudaGetErrorString(e));
        tfbs_shapes_destroy(s);
        return e == cudaErrorMemoryAllocation ? TFBS_ERR_NOMEM : TFBS_ERR_CUDA;
    }
    *out = s;
    return TFBS_OK;
}

int tfbs_shapes_destroy(tfbs_shapes *s)
{
    if (!s) return TFBS_OK;
    if (s->ctx) cudaSetDevice(s->ctx->device);
    if (s->tables) cudaFree(s->tables);
    if (s->kinds_dev) cudaFree(s->kinds_dev);
    if (s->c16_words) cudaFree(s->c16_words);
    delete s;
    return TFBS_OK;
}

int tfbs_shape_tracks(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_shapes *shapes,
                      const int64_t *rec_start, const int64_t *rec_len, int64_t S, float *out_host)
{
    TFBS_REQUIRE(ctx && seq && shapes && S >= 0, "tfbs_shape_tracks: bad argument");
    TFBS_REQUIRE(seq->ctx == ctx && shapes->ctx == ctx, "tfbs_shape_tracks: handle from another context");
    TFBS_REQUIRE(S == 0 || (rec_start && rec_len), "tfbs_shape_tracks: null record arrays");
    const int64_t n = seq->n;
    if (n == 0) return TFBS_OK;
    for (int64_t i = 0; i < S; i++)
        TFBS_REQUIRE(rec_start[i] >= 0 && rec_len[i] >= 0 && rec_start[i] + rec_len[i] <= n,
                     "tfbs_shape_tracks: record %lld out of range", (long long)i);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int T = shapes->T;
    const int64_t pitch = ((n + kShapeTile - 1) / kShapeTile) * kShapeTile;
    int rc = tfbs_scratch_reserve(ctx->tracks, (size_t)T * pitch * sizeof(float));
    if (rc) return rc;
    // record table for the edge fix-up (one implicit record [0, n) when none is given)
    const int64_t one_start = 0, one_len = n;
    if (S == 0) {
        rec_start = &one_start;
        rec_len = &one_len;
        S = 1;
    }
    rc = tfbs_scratch_reserve(ctx->sites, (size_t)S * 16);
    if (rc) return rc;
    int64_t *d_rs = (int64_t *)ctx->sites.ptr, *d_rl = d_rs + S;
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_rs, rec_start, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_rl, rec_len, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));  // rec_start may point at our stack

    TrackParams P;
    P.packed = seq->packed;
    P.mask = seq->mask;
    P.n_words = seq->n_words;
    P.n_mask = seq->n_mask;
    P.pitch = pitch;
    P.tables = shapes->tables;
    P.T = T;
    P.step_mask = 0;
    for (int t = 0; t < T; t++)
        if (shapes->kinds[t] == TFBS_SHAPE_PER_STEP) P.step_mask |= 1u << t;
    P.out = (float *)ctx->tracks.ptr;
    const int64_t grid = std::min<int64_t>(pitch / (kShapeThreads * kShapePosPerThread), (int64_t)ctx->sm_count * 8);
    tfbs_begin(ctx);
    if (shapes->c16_words) {
        Packed16Params Q;
        Q.packed = seq->packed;
        Q.mask = seq->mask;
        Q.n_words = seq->n_words;
        Q.n_mask = seq->n_mask;
        Q.pitch = pitch;
        Q.words = shapes->c16_words;
        Q.NW = (int32_t)shapes->c16_kind.size();
        for (int w = 0; w < kMaxTracks; w++) {
            Q.kind[w] = w < Q.NW ? shapes->c16_kind[w] : 0;
            Q.off_a[w] = w < Q.NW ? (int64_t)shapes->c16_row_a[w] * pitch : 0;
            Q.off_b[w] = (w < Q.NW && shapes->c16_row_b[w] >= 0) ? (int64_t)shapes->c16_row_b[w] * pitch : -1;
        }
        Q.out = P.out;
        shape_tracks_c16_kernel<<<(unsigned)grid, kShapeThreads, (size_t)Q.NW * 4096, ctx->stream>>>(Q);
    } else {
        const size_t smem = (size_t)T * 2048 * sizeof(float);
        TFBS_CHECK_CUDA(cudaFuncSetAttribute(shape_tracks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        shape_tracks_kernel<<<(unsigned)grid, kShapeThreads, smem, ctx->stream>>>(P);
    }
    tfbs_launched(ctx);
    {
        const int64_t total = S * 2 * T;
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
        shape_edge_fix_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            seq->packed, seq->mask, shapes->tables, T, shapes->kinds_dev, d_rs, d_rl, S, pitch, P.out);
        tfbs_launched(ctx);
    }
    rc = tfbs_end(ctx);
    if (rc) return rc;
    if (out_host)
        TFBS_CHECK_CUDA(cudaMemcpy2D(out_host, (size_t)n * 4, ctx->tracks.ptr, (size_t)pitch * 4,
                                     (size_t)n * 4, T, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}

int tfbs_shape_sites(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_shapes *shapes,
                     const int64_t *rec_start, const int64_t *rec_len, const int64_t *start,
                     const int64_t *end, const int8_t *strand, int64_t S, int32_t W, int32_t mode,
                     int32_t trailing_empty, float *out_host)
{
    TFBS_REQUIRE(ctx && seq && shapes && S >= 0 && W > 0, "tfbs_shape_sites: bad argument");
    TFBS_REQUIRE(seq->ctx == ctx && shapes->ctx == ctx, "tfbs_shape_sites: handle from another context");
    TFBS_REQUIRE(mode == 0 || mode == 1, "tfbs_shape_sites: mode %d", mode);
    if (S == 0) return TFBS_OK;
    TFBS_REQUIRE(rec_start && rec_len && out_host, "tfbs_shape_sites: null argument");
    TFBS_REQUIRE(mode == 1 || (start && end && strand), "tfbs_shape_sites: mode 0 needs start/end/strand");
    for (int64_t i = 0; i < S; i++)
        TFBS_REQUIRE(rec_start[i] >= 0 && rec_len[i] >= 0 && rec_start[i] + rec_len[i] <= seq->n,
                     "tfbs_shape_sites: record of site %lld out of range", (long long)i);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int T = shapes->T;
    const size_t out_bytes = (size_t)S * T * W * sizeof(float);
    const size_t arg_bytes = (size_t)S * (8 * 4 + 8);  // 4 int64 arrays + strand (padded)
    int rc = tfbs_scratch_reserve(ctx->sites, arg_bytes + out_bytes + 256);
    if (rc) return rc;
    char *base = (char *)ctx->sites.ptr;
    int64_t *d_rs = (int64_t *)base, *d_rl = d_rs + S, *d_st = d_rl + S, *d_en = d_st + S;
    int8_t *d_sd = (int8_t *)(d_en + S);
    float *d_out = (float *)(base + ((arg_bytes + 255) / 256) * 256);
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_rs, rec_start, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_rl, rec_len, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (mode == 0) {
        TFBS_CHECK_CUDA(cudaMemcpyAsync(d_st, start, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
        TFBS_CHECK_CUDA(cudaMemcpyAsync(d_en, end, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
        TFBS_CHECK_CUDA(cudaMemcpyAsync(d_sd, strand, (size_t)S, cudaMemcpyHostToDevice, ctx->stream));
    }
    tfbs_begin(ctx);
    const int64_t total = S * T * W;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    shape_sites_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
        seq->packed, seq->mask, shapes->tables, T, shapes->kinds_dev, d_rs, d_rl, d_st, d_en, d_sd, S, W,
        mode, trailing_empty, d_out);
    tfbs_launched(ctx);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaMemcpy(out_host, d_out, out_bytes, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}

int tfbs_track_sites(tfbs_ctx *ctx, const float *values_host, int64_t n_values,
                     const int64_t *rec_off, const int64_t *rec_ntok, const int64_t *start,
                     const int64_t *end, const int8_t *strand, int64_t S, int32_t W, float *out_host)
{
    TFBS_REQUIRE(ctx && S >= 0 && W > 0 && n_values >= 0, "tfbs_track_sites: bad argument");
    if (S == 0) return TFBS_OK;
    TFBS_REQUIRE(values_host && rec_off && rec_ntok && start && end && strand && out_host,
                 "tfbs_track_sites: null argument");
    for (int64_t i = 0; i < S; i++)
        TFBS_REQUIRE(rec_off[i] >= 0 && rec_ntok[i] >= 0 && rec_off[i] + rec_ntok[i] <= n_values,
                     "tfbs_track_sites: token list of site %lld out of range", (long long)i);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const size_t val_bytes = (((size_t)n_values * 4 + 255) / 256) * 256;
    const size_t arg_bytes = (((size_t)S * 33 + 255) / 256) * 256;
    const size_t out_bytes = (size_t)S * W * sizeof(float);
    int rc = tfbs_scratch_reserve(ctx->tracks, val_bytes + 256);
    if (rc) return rc;
    rc = tfbs_scratch_reserve(ctx->sites, arg_bytes + out_bytes + 256);
    if (rc) return rc;
    float *d_val = (float *)ctx->tracks.ptr;
    char *base = (char *)ctx->sites.ptr;
    int64_t *d_off = (int64_t *)base, *d_nt = d_off + S, *d_st = d_nt + S, *d_en = d_st + S;
    int8_t *d_sd = (int8_t *)(d_en + S);
    float *d_out = (float *)(base + arg_bytes);
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_val, values_host, (size_t)n_values * 4, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_off, rec_off, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_nt, rec_ntok, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_st, start, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_en, end, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_sd, strand, (size_t)S, cudaMemcpyHostToDevice, ctx->stream));
    tfbs_begin(ctx);
    const int64_t total = S * W;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    track_sites_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_val, d_off, d_nt, d_st, d_en, d_sd, S, W, d_out);
    tfbs_launched(ctx);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaMemcpy(out_host, d_out, out_bytes, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}

}  // extern "C"
