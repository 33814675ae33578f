// encode.cu — (1) sequence encoding: ASCII -> 2-bit words + invalid-bit mask, base counts,
// per-record G+C counts and the device-side synthetic genome generator.
//
// Replaces the per-character Python string work of the reference (utils.py:97-105
// reverse_complement, utils.py:224-238 calculate_gc_content) and the byte switch at the top of
// Biopython's _pwm.c inner loop.  HBM-bound: 1 B read + 0.375 B written per base.
#include "common.cuh"
#include "swar.h"

namespace {

constexpr int kEncodeThreads = 256;
constexpr int kBasesPerWarpIter = 2048;  // 4 x (32 lanes x 16 bytes)

__device__ __forceinline__ uint4 ld_stream_v4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint4 ld_bytes_guarded(const uint8_t *__restrict__ a, int64_t idx,
                                                  int64_t n)
{
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int64_t i = idx + k;
        uint32_t b = (i < n) ? (uint32_t)a[i] : 0u;
        w[k >> 2] |= b << (8 * (k & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// One lane converts 16 bases per step; a warp covers 2048 consecutive bases per iteration so
// that the 128-bit loads, the packed-word stores (128 B / warp) and the mask stores (64 B /
// warp) are all fully coalesced.  n_pad is a multiple of 2048, so every lane is always active.
__global__ void __launch_bounds__(kEncodeThreads)
encode_kernel(const uint8_t *__restrict__ ascii, int64_t n, int64_t n_pad,
              uint32_t *__restrict__ packed, uint32_t *__restrict__ mask)
{
    // `ascii`, `packed` and `mask` point at the first base of the range being encoded (a multiple
    // of 2048 bases into the buffer); n = valid bytes of the range, n_pad = bases it must cover.
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * (kEncodeThreads / 32) + (threadIdx.x >> 5);
    const int64_t warp_stride = (int64_t)gridDim.x * (kEncodeThreads / 32);
    const bool aligned = ((reinterpret_cast<uintptr_t>(ascii) & 15) == 0);
    const int64_t n_chunks = n_pad / kBasesPerWarpIter;

    for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warp_stride) {
        const int64_t base = chunk * kBasesPerWarpIter + lane * 16;
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t idx = base + j * 512;
            if (aligned && idx + 16 <= n) v[j] = ld_stream_v4(ascii + idx);
            else if (idx < n) v[j] = ld_bytes_guarded(ascii, idx, n);
            else v[j] = make_uint4(0, 0, 0, 0);
        }
        // fast path: assume every byte is ACGTacgt (true for almost every chunk of a genome);
        // the per-byte validity flags are only OR-reduced here
        uint32_t word[4], bad = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t d0, d1, d2, d3;
            const uint32_t g0 = tfbs_gather_top(tfbs_swar_codes(v[j].x, &d0));
            const uint32_t g1 = tfbs_gather_top(tfbs_swar_codes(v[j].y, &d1));
            const uint32_t g2 = tfbs_gather_top(tfbs_swar_codes(v[j].z, &d2));
            const uint32_t g3 = tfbs_gather_top(tfbs_swar_codes(v[j].w, &d3));
            // top bytes of g0..g3 -> bytes 0..3 of the packed word
            word[j] = __byte_perm(__byte_perm(g0, g1, 0x0073), __byte_perm(g2, g3, 0x0073), 0x5410);
            bad |= d0 | d1 | d2 | d3;
        }
        if (!__any_sync(0xFFFFFFFFu, bad != 0u)) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t idx = base + j * 512;
                packed[idx >> 4] = word[j];
                if ((lane & 1) == 0) mask[idx >> 5] = 0u;
            }
        } else {
            // some byte of this 2048-base chunk is not ACGT (N-run, IUPAC code, separator, padding)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t idx = base + j * 512;
                uint32_t v0, v1, v2, v3;
                uint32_t wd = tfbs_swar4(v[j].x, &v0);
                wd |= tfbs_swar4(v[j].y, &v1) << 8;
                wd |= tfbs_swar4(v[j].z, &v2) << 16;
                wd |= tfbs_swar4(v[j].w, &v3) << 24;
                const uint32_t valid16 = v0 | (v1 << 4) | (v2 << 8) | (v3 << 12);
                const uint32_t inv16 = ~valid16 & 0xFFFFu;
                packed[idx >> 4] = wd;
                const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, inv16, 1);
                if ((lane & 1) == 0) mask[idx >> 5] = inv16 | (other << 16);
            }
        }
    }
}

// counts[0..3] = A C G T, counts[4] = invalid, over bases [0, n).
__global__ void __launch_bounds__(256)
base_counts_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask, int64_t n,
                   unsigned long long *__restrict__ counts)
{
    unsigned long long c1 = 0, c2 = 0, c3 = 0, cinv = 0, ctot = 0;
    const int64_t n_mask = (n + 31) / 32;
    for (int64_t mw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; mw < n_mask;
         mw += (int64_t)gridDim.x * blockDim.x) {
        uint32_t inv = mask[mw];
        int64_t rem = n - mw * 32;
        uint32_t in_range = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
        inv &= in_range;
        cinv += __popc(inv);
        ctot += __popc(in_range);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // positions past n and invalid bases are stored as code 0, so C/G/T need no masking
            uint32_t p = packed[mw * 2 + h];
            uint32_t lo = p & 0x55555555u, hi = (p >> 1) & 0x55555555u;
            c1 += __popc(lo & ~hi);
            c2 += __popc(hi & ~lo);
            c3 += __popc(lo & hi);
        }
    }
    __shared__ unsigned long long red[5];
    if (threadIdx.x < 5) red[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long vals[5] = {ctot - cinv - c1 - c2 - c3, c1, c2, c3, cinv};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned long long v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 5) atomicAdd(&counts[threadIdx.x], red[threadIdx.x]);
}

// One warp per record: number of C or G bases in [start, start+len).
__global__ void __launch_bounds__(256)
gc_counts_kernel(const uint32_t *__restrict__ packed, const int64_t *__restrict__ start,
                 const int64_t *__restrict__ len, int64_t S, long long *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    for (int64_t rec = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rec < S;
         rec += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t s = start[rec], e = s + len[rec];
        long long cnt = 0;
        for (int64_t w = (s >> 4) + lane; w <= ((e - 1) >> 4) && e > s; w += 32) {
            uint32_t p = __ldg(packed + w);
            uint32_t gc = (p ^ (p >> 1)) & 0x55555555u;  // code 1 (C) or 2 (G): bits differ
            const int64_t wb = w << 4;                     // first base of this word
            if (wb < s) gc &= ~0u << (2 * (s - wb));
            if (wb + 16 > e) gc &= (e - wb >= 16) ? ~0u : ((1u << (2 * (e - wb))) - 1u);
            cnt += __popc(gc);
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        if (lane == 0) out[rec] = cnt;
    }
}

// ---- synthetic genome (position-addressable; mirrors discover_tfbs_b200/synth.py) -----------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_mix(uint64_t seed, uint64_t tag, uint64_t i)
{
    return splitmix64(splitmix64(seed + tag) ^ i);
}

__global__ void __launch_bounds__(256)
synth_kernel(uint64_t seed, int64_t start, int64_t n, uint64_t t_c, uint64_t t_g, uint64_t t_a,
             int with_n, int with_lower, uint8_t *__restrict__ out)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 16 < n;
         q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int64_t i = start + q * 16 + k;
            const uint64_t r = synth_mix(seed, 1, (uint64_t)i) >> 11;  // 53 bits
            uint32_t ch = r < t_c ? 'C' : r < t_g ? 'G' : r < t_a ? 'A' : 'T';
            if (with_lower) {
                const uint64_t rl = synth_mix(seed, 3, (uint64_t)(i >> 6));
                if ((rl & 0xFFFF) < 3277) ch |= 0x20;
            }
            if (with_n) {
                const uint64_t rn = synth_mix(seed, 2, (uint64_t)(i >> 14));
                if ((rn & 0xFFFF) < 213) {
                    const int64_t off = i & 16383;
                    const int64_t s = (int64_t)((rn >> 16) & 16383);
                    const int64_t l = 100 + (int64_t)((rn >> 32) % 9901);
                    if (off >= s && off < s + l) ch = 'N';
                }
            }
            w[k >> 2] |= ch << (8 * (k & 3));
        }
        if (q * 16 + 16 <= n) {
            reinterpret_cast<uint4 *>(out)[q] = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int k = 0; q * 16 + k < n; k++) out[q * 16 + k] = (w[k >> 2] >> (8 * (k & 3))) & 0xFF;
        }
    }
}

int alloc_seq(tfbs_ctx *ctx, int64_t n, tfbs_seq **out)
{
    tfbs_seq *s = new tfbs_seq();
    s->ctx = ctx;
    s->n = n;
    const int64_t n_pad = ((n + 8191) / 8192) * 8192 + 5 * 8192;  // tiles of every kernel may over-read (hits: 32 256 + 128)
    s->n_words = n_pad / 16;
    s->n_mask = n_pad / 32;
    // 64-word front pad (bases "-1024..-1": code 0, invalid) lets the shape kernels read the window
    // left of position 0 without bounds checks
    constexpr int kFrontPad = 64;
    cudaError_t e = cudaMalloc(&s->packed_base, (size_t)(s->n_words + kFrontPad) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s->mask_base, (size_t)(s->n_mask + kFrontPad) * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->packed_base, 0, kFrontPad * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->mask_base, 0xFF, kFrontPad * 4, ctx->stream);
    if (e == cudaSuccess) {
        s->packed = s->packed_base + kFrontPad;
        s->mask = s->mask_base + kFrontPad;
    }
    if (e != cudaSuccess) {
        tfbs_set_error("cudaMalloc for a %lld-base sequence failed: %s", (long long)n,
                       cudaGetErrorString(e));
        if (s->packed_base) cudaFree(s->packed_base);
        if (s->mask_base) cudaFree(s->mask_base);
        delete s;
        return TFBS_ERR_NOMEM;
    }
    *out = s;
    return TFBS_OK;
}

}  // namespace

// Encode `valid` bytes at ascii_dev into bases [base, base + cover) of seq (base and cover are
// multiples of 2048; bytes past `valid` are marked invalid).  Used by the chunk-pipelined host path.
int tfbs_encode_range(tfbs_ctx *ctx, const void *ascii_dev, int64_t valid, tfbs_seq *s, int64_t base,
                      int64_t cover, cudaStream_t stream)
{
    const int64_t chunks = cover / kBasesPerWarpIter;
    const int64_t blocks = (chunks + (kEncodeThreads / 32) - 1) / (kEncodeThreads / 32);
    if (blocks > 0)
        encode_kernel<<<(unsigned)blocks, kEncodeThreads, 0, stream>>>((const uint8_t *)ascii_dev, valid, cover,
                                                                      s->packed + (base >> 4), s->mask + (base >> 5));
    tfbs_launched(ctx);
    return TFBS_OK;
}

namespace {

int launch_encode(tfbs_ctx *ctx, const void *ascii_dev, int64_t n, tfbs_seq *s)
{
    const int64_t n_pad = s->n_words * 16;
    const int64_t chunks = n_pad / kBasesPerWarpIter;
    // one 2048-base chunk per warp: the block scheduler balances the tail better than a
    // grid-stride loop with ~5 chunks per warp (17 % imbalance at 100 Mb)
    int64_t blocks = (chunks + (kEncodeThreads / 32) - 1) / (kEncodeThreads / 32);
    const int64_t max_blocks = 1 << 22;
    if (blocks > max_blocks) blocks = max_blocks;
    encode_kernel<<<(unsigned)blocks, kEncodeThreads, 0, ctx->stream>>>(
        (const uint8_t *)ascii_dev, n, n_pad, s->packed, s->mask);
    tfbs_launched(ctx);
    return TFBS_OK;
}

}  // namespace

extern "C" {

int tfbs_encode_device_into(tfbs_ctx *ctx, const void *ascii_dev, int64_t n, tfbs_seq *seq)
{
    TFBS_REQUIRE(ctx && seq && (ascii_dev || n == 0), "tfbs_encode_device_into: null argument");
    TFBS_REQUIRE(n == seq->n, "tfbs_encode_device_into: length %lld != handle length %lld",
                 (long long)n, (long long)seq->n);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    tfbs_begin(ctx);
    launch_encode(ctx, ascii_dev, n, seq);
    return tfbs_end(ctx);
}

int tfbs_encode_device(tfbs_ctx *ctx, const void *ascii_dev, int64_t n, tfbs_seq **out)
{
    TFBS_REQUIRE(ctx && out && n >= 0 && (ascii_dev || n == 0), "tfbs_encode_device: bad argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    tfbs_seq *s = nullptr;
    int rc = alloc_seq(ctx, n, &s);
    if (rc) return rc;
    rc = tfbs_encode_device_into(ctx, ascii_dev, n, s);
    if (rc) {
        tfbs_seq_destroy(s);
        return rc;
    }
    *out = s;
    return TFBS_OK;
}

int tfbs_seq_create(tfbs_ctx *ctx, int64_t n, tfbs_seq **out)
{
    TFBS_REQUIRE(ctx && out && n >= 0, "tfbs_seq_create: bad argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    tfbs_seq *s = nullptr;
    int rc = alloc_seq(ctx, n, &s);
    if (rc) return rc;
    // until something is encoded into it the handle holds n invalid bases
    cudaError_t e = cudaMemsetAsync(s->packed, 0, (size_t)s->n_words * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->mask, 0xFF, (size_t)s->n_mask * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        tfbs_set_error("tfbs_seq_create: %s", cudaGetErrorString(e));
        tfbs_seq_destroy(s);
        return TFBS_ERR_CUDA;
    }
    *out = s;
    return TFBS_OK;
}

int tfbs_encode_into(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq *seq)
{
    TFBS_REQUIRE(ctx && seq && (ascii_host || n == 0), "tfbs_encode_into: null argument");
    TFBS_REQUIRE(n == seq->n, "tfbs_encode_into: length %lld != handle length %lld", (long long)n,
                 (long long)seq->n);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    if (seq->staging_bytes < (size_t)n) {
        if (seq->staging) cudaFree(seq->staging);
        seq->staging = nullptr;
        seq->staging_bytes = 0;
        TFBS_CHECK_CUDA(cudaMalloc(&seq->staging, (size_t)(n > 0 ? n : 1)));
        seq->staging_bytes = (size_t)n;
    }
    if (n > 0)
        TFBS_CHECK_CUDA(cudaMemcpyAsync(seq->staging, ascii_host, (size_t)n, cudaMemcpyHostToDevice,
                                        ctx->stream));
    tfbs_begin(ctx);
    launch_encode(ctx, seq->staging, n, seq);
    return tfbs_end(ctx);
}

int tfbs_encode(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq **out)
{
    TFBS_REQUIRE(ctx && out && n >= 0 && (ascii_host || n == 0), "tfbs_encode: bad argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    tfbs_seq *s = nullptr;
    int rc = alloc_seq(ctx, n, &s);
    if (rc) return rc;
    rc = tfbs_encode_into(ctx, ascii_host, n, s);
    if (rc) {
        tfbs_seq_destroy(s);
        return rc;
    }
    *out = s;
    return TFBS_OK;
}

int tfbs_seq_destroy(tfbs_seq *seq)
{
    if (!seq) return TFBS_OK;
    if (seq->ctx) cudaSetDevice(seq->ctx->device);
    if (seq->packed_base) cudaFree(seq->packed_base);
    if (seq->mask_base) cudaFree(seq->mask_base);
    if (seq->staging) cudaFree(seq->staging);
    delete seq;
    return TFBS_OK;
}

int tfbs_seq_length(const tfbs_seq *seq, int64_t *n)
{
    TFBS_REQUIRE(seq && n, "tfbs_seq_length: null argument");
    *n = seq->n;
    return TFBS_OK;
}

int tfbs_seq_download(tfbs_ctx *ctx, const tfbs_seq *seq, uint32_t *packed, uint32_t *mask)
{
    TFBS_REQUIRE(ctx && seq, "tfbs_seq_download: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (packed)
        TFBS_CHECK_CUDA(cudaMemcpy(packed, seq->packed, (size_t)((seq->n + 15) / 16) * 4,
                                   cudaMemcpyDeviceToHost));
    if (mask)
        TFBS_CHECK_CUDA(
            cudaMemcpy(mask, seq->mask, (size_t)((seq->n + 31) / 32) * 4, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}

int tfbs_seq_base_counts(tfbs_ctx *ctx, const tfbs_seq *seq, int64_t counts[5])
{
    TFBS_REQUIRE(ctx && seq && counts, "tfbs_seq_base_counts: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    int rc = tfbs_scratch_reserve(ctx->misc, 64);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaMemsetAsync(ctx->misc.ptr, 0, 5 * sizeof(unsigned long long), ctx->stream));
    tfbs_begin(ctx);
    if (seq->n > 0) {
        int64_t blocks = ((seq->n + 31) / 32 + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
        base_counts_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            seq->packed, seq->mask, seq->n, (unsigned long long *)ctx->misc.ptr);
        tfbs_launched(ctx);
    }
    rc = tfbs_end(ctx);
    if (rc) return rc;
    unsigned long long h[5];
    TFBS_CHECK_CUDA(cudaMemcpy(h, ctx->misc.ptr, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 5; k++) counts[k] = (int64_t)h[k];
    return TFBS_OK;
}

int tfbs_gc_counts(tfbs_ctx *ctx, const tfbs_seq *seq, const int64_t *start, const int64_t *len,
                   int64_t S, int64_t *gc_out)
{
    TFBS_REQUIRE(ctx && seq && (S == 0 || (start && len && gc_out)), "tfbs_gc_counts: null argument");
    if (S == 0) return TFBS_OK;
    for (int64_t i = 0; i < S; i++)
        TFBS_REQUIRE(start[i] >= 0 && len[i] >= 0 && start[i] + len[i] <= seq->n,
                     "tfbs_gc_counts: record %lld out of range", (long long)i);
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    int rc = tfbs_scratch_reserve(ctx->sites, (size_t)S * 24);
    if (rc) return rc;
    int64_t *d_start = (int64_t *)ctx->sites.ptr, *d_len = d_start + S;
    long long *d_out = (long long *)(d_len + S);
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_start, start, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    TFBS_CHECK_CUDA(cudaMemcpyAsync(d_len, len, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    tfbs_begin(ctx);
    int64_t blocks = (S + 7) / 8;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    gc_counts_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(seq->packed, d_start, d_len, S, d_out);
    tfbs_launched(ctx);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaMemcpy(gc_out, d_out, (size_t)S * 8, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}

int tfbs_synth_ascii(tfbs_ctx *ctx, uint64_t seed, int64_t start, int64_t n, double gc,
                     int32_t with_n_runs, int32_t with_lowercase, void **ascii_dev_out)
{
    TFBS_REQUIRE(ctx && ascii_dev_out && n >= 0 && start >= 0 && gc >= 0.0 && gc <= 1.0,
                 "tfbs_synth_ascii: bad argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    void *buf = nullptr;
    TFBS_CHECK_CUDA(cudaMalloc(&buf, (size_t)(n > 0 ? ((n + 15) / 16) * 16 : 16)));
    // thresholds on the 53-bit uniform integer, computed exactly as synth.py does
    const double two53 = 9007199254740992.0;
    const uint64_t t_c = (uint64_t)(gc / 2.0 * two53);
    const uint64_t t_g = (uint64_t)(gc * two53);
    const uint64_t t_a = (uint64_t)((gc + (1.0 - gc) / 2.0) * two53);
    tfbs_begin(ctx);
    if (n > 0) {
        int64_t blocks = ((n + 15) / 16 + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
        synth_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(seed, start, n, t_c, t_g, t_a,
                                                                with_n_runs, with_lowercase,
                                                                (uint8_t *)buf);
        tfbs_launched(ctx);
    }
    int rc = tfbs_end(ctx);
    if (rc) {
        cudaFree(buf);
        return rc;
    }
    *ascii_dev_out = buf;
    return TFBS_OK;
}

// Host-side check of the SWAR bit tricks (no device needed; exercised by the CPU test-suite).
int tfbs_debug_swar4(uint32_t w, uint32_t *codes8, uint32_t *valid4)
{
    if (!codes8 || !valid4) return TFBS_ERR_INVALID;
    *codes8 = tfbs_swar4(w, valid4);
    return TFBS_OK;
}

}  // extern "C"
