// similarity.cu — all-pairs MinHash-Jaccard and Poisson (PAC) similarity sums  (SURVEY.md §8(f)
// rank 1: the reference spends 98 % of its own feature-build time here).
//
// Replaces the nested Python loops of utils.py:1043-1058 (_minhash_similarity, _pac) over
// utils.py:1243-1255 (jaccard_similarity) and utils.py:892-949 (pac).  The host prepares, per
// sequence, the sorted unique k-mer codes with multiplicities and the sorted unique canonical
// k-mer hashes (MurmurHash3_x86_32, seed 39 — utils.py:1259-1269); this kernel evaluates every
// (candidate, true) pair: one warp per candidate, lanes stride over the true sequences, double
// precision throughout (the columns are float64 in the reference's DataFrame).
#include "common.cuh"

namespace {

struct SimSet {
    const uint32_t *words;   // sorted unique k-mer codes, concatenated
    const uint16_t *counts;  // multiplicity of every unique k-mer
    const int32_t *woff;     // [N+1]
    const int32_t *nk;       // [N] total k-mers (with multiplicity)
    const uint32_t *hash;    // sorted unique canonical hashes, concatenated
    const int32_t *hoff;     // [N+1]
};

// P(X >= c) for X ~ Poisson(mu), c >= 1:  1 - exp(-mu) * sum_{i<c} mu^i / i!
// (the reference composes 1 - scipy.stats.poisson.cdf(c - 1, mu), utils.py:881-888,929)
__device__ __forceinline__ double poisson_sf(int c, double mu)
{
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < c; i++) {
        term *= mu / (double)i;
        sum += term;
    }
    return 1.0 - exp(-mu) * sum;
}

__global__ void __launch_bounds__(256)
similarity_kernel(const SimSet C, const SimSet T, int64_t S, int64_t U, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < S;
         s += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int cw0 = __ldg(C.woff + s), cw1 = __ldg(C.woff + s + 1);
        const int ch0 = __ldg(C.hoff + s), ch1 = __ldg(C.hoff + s + 1);
        const int nB = __ldg(C.nk + s);
        double jac = 0.0, pas = 0.0, pps = 0.0;
        for (int64_t u = lane; u < U; u += 32) {
            // ---- Jaccard over canonical hashes: sorted-list merge -----------------------------------
            const int th0 = __ldg(T.hoff + u), th1 = __ldg(T.hoff + u + 1);
            int i = th0, j = ch0, inter = 0;
            while (i < th1 && j < ch1) {
                const uint32_t a = __ldg(T.hash + i), b = __ldg(C.hash + j);
                inter += a == b;
                i += a <= b;
                j += b <= a;
            }
            const int uni = (th1 - th0) + (ch1 - ch0) - inter;
            jac += (double)inter / (double)uni;
            // ---- PAC: loop over the k-mers of the TRUE sequence (seqA in utils.py:892) ----------------
            const int tw0 = __ldg(T.woff + u), tw1 = __ldg(T.woff + u + 1);
            const int nA = __ldg(T.nk + u);
            double prod = 1.0, sim = 0.0;
            int jb = cw0;
            for (int ia = tw0; ia < tw1; ia++) {
                const uint32_t w = __ldg(T.words + ia);
                const int ca = __ldg(T.counts + ia);
                while (jb < cw1 && __ldg(C.words + jb) < w) jb++;
                const int cb = (jb < cw1 && __ldg(C.words + jb) == w) ? (int)__ldg(C.counts + jb) : 0;
                const int c = ca < cb ? ca : cb;
                double prob = 1.0;
                if (c > 0) {
                    const double mwA = ((double)ca / (double)nA) * (double)nA;
                    const double mwB = ((double)cb / (double)nB) * (double)nB;
                    prob = poisson_sf(c, mwA) * poisson_sf(c, mwB);
                }
                // the reference visits every occurrence of the word in seqA
                for (int r = 0; r < ca; r++) {
                    prod *= prob;
                    sim += 1.0 - prob;
                }
            }
            pas += sim / (double)nA;
            pps += 1.0 - pow(prod, 1.0 / (double)nA);
        }
        for (int o = 16; o > 0; o >>= 1) {
            jac += __shfl_xor_sync(0xFFFFFFFFu, jac, o);
            pas += __shfl_xor_sync(0xFFFFFFFFu, pas, o);
            pps += __shfl_xor_sync(0xFFFFFFFFu, pps, o);
        }
        if (lane == 0) {
            out[s * 3 + 0] = jac;
            out[s * 3 + 1] = pas;
            out[s * 3 + 2] = pps;
        }
    }
}

}  // namespace

extern "C" int tfbs_similarity_sums(tfbs_ctx *ctx, const uint32_t *c_words, const uint16_t *c_counts,
                                    const int32_t *c_woff, const int32_t *c_nk, const uint32_t *c_hash,
                                    const int32_t *c_hoff, int64_t S, const uint32_t *t_words,
                                    const uint16_t *t_counts, const int32_t *t_woff, const int32_t *t_nk,
                                    const uint32_t *t_hash, const int32_t *t_hoff, int64_t U, double *out_host)
{
    TFBS_REQUIRE(ctx && S >= 0 && U >= 0, "tfbs_similarity_sums: bad argument");
    if (S == 0) return TFBS_OK;
    TFBS_REQUIRE(c_words && c_counts && c_woff && c_nk && c_hash && c_hoff && out_host,
                 "tfbs_similarity_sums: null candidate argument");
    TFBS_REQUIRE(U == 0 || (t_words && t_counts && t_woff && t_nk && t_hash && t_hoff),
                 "tfbs_similarity_sums: null true-set argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int64_t cw = c_woff[S], chn = c_hoff[S], tw = U ? t_woff[U] : 0, thn = U ? t_hoff[U] : 0;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t need = al(cw * 4) + al(cw * 2) + al((S + 1) * 4) + al(S * 4) + al(chn * 4) + al((S + 1) * 4) +
                        al(tw * 4) + al(tw * 2) + al((U + 1) * 4) + al(U * 4) + al(thn * 4) + al((U + 1) * 4) +
                        al(S * 3 * 8) + 256;
    int rc = tfbs_scratch_reserve(ctx->sites, need);
    if (rc) return rc;
    char *p = (char *)ctx->sites.ptr;
    auto put = [&](const void *src, size_t bytes) -> void * {
        void *d = p;
        if (bytes) cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
        p += al(bytes);
        return d;
    };
    SimSet C, T;
    C.words = (const uint32_t *)put(c_words, cw * 4);
    C.counts = (const uint16_t *)put(c_counts, cw * 2);
    C.woff = (const int32_t *)put(c_woff, (S + 1) * 4);
    C.nk = (const int32_t *)put(c_nk, S * 4);
    C.hash = (const uint32_t *)put(c_hash, chn * 4);
    C.hoff = (const int32_t *)put(c_hoff, (S + 1) * 4);
    T.words = (const uint32_t *)put(t_words, tw * 4);
    T.counts = (const uint16_t *)put(t_counts, tw * 2);
    T.woff = (const int32_t *)put(t_woff, (U + 1) * 4);
    T.nk = (const int32_t *)put(t_nk, U * 4);
    T.hash = (const uint32_t *)put(t_hash, thn * 4);
    T.hoff = (const int32_t *)put(t_hoff, (U + 1) * 4);
    double *d_out = (double *)p;
    TFBS_CHECK_CUDA(cudaGetLastError());
    tfbs_begin(ctx);
    int64_t blocks = (S + 7) / 8;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    similarity_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(C, T, S, U, d_out);
    tfbs_launched(ctx);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaMemcpy(out_host, d_out, (size_t)S * 3 * 8, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}
