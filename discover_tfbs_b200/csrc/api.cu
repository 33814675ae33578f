// api.cu — context, error reporting and memory plumbing of libtfbs_cuda.so.
#include <cstring>

#include "common.cuh"

static thread_local char g_err[512] = "";

void tfbs_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int tfbs_scratch_reserve(tfbs_scratch &s, size_t bytes)
{
    if (s.bytes >= bytes && s.ptr) return TFBS_OK;
    if (s.ptr) cudaFree(s.ptr);
    s.ptr = nullptr;
    s.bytes = 0;
    cudaError_t e = cudaMalloc(&s.ptr, bytes);
    if (e != cudaSuccess) {
        tfbs_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        s.ptr = nullptr;
        cudaGetLastError();
        return TFBS_ERR_NOMEM;
    }
    s.bytes = bytes;
    return TFBS_OK;
}

namespace {
__global__ void __launch_bounds__(256) flush_kernel(uint4 *p, size_t n16, uint32_t v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(v, v, v, v);
}
void free_scratch(tfbs_scratch &s)
{
    if (s.ptr) cudaFree(s.ptr);
    s.ptr = nullptr;
    s.bytes = 0;
}
}  // namespace

extern "C" {

const char *tfbs_last_error(void) { return g_err; }

const char *tfbs_version(void) { return "discover_tfbs_b200 libtfbs_cuda 0.1 (sm_100a)"; }

int tfbs_device_count(int *n)
{
    TFBS_REQUIRE(n, "tfbs_device_count: null argument");
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) {
        *n = 0;
        tfbs_set_error("no CUDA device: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return TFBS_ERR_NO_DEVICE;
    }
    return TFBS_OK;
}

int tfbs_ctx_create(int device, tfbs_ctx **out)
{
    TFBS_REQUIRE(out, "tfbs_ctx_create: null argument");
    int n = 0;
    int rc = tfbs_device_count(&n);
    if (rc) return rc;
    if (n == 0) {
        tfbs_set_error("no CUDA device visible; libtfbs_cuda has no CPU fallback");
        return TFBS_ERR_NO_DEVICE;
    }
    TFBS_REQUIRE(device >= 0 && device < n, "tfbs_ctx_create: device %d out of range (0..%d)", device, n - 1);
    TFBS_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TFBS_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        tfbs_set_error("device %d is sm_%d%d; libtfbs_cuda is built for sm_100a (B200) only", device,
                       prop.major, prop.minor);
        return TFBS_ERR_NO_DEVICE;
    }
    tfbs_ctx *c = new tfbs_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_mid);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_direct);
    if (e == cudaSuccess) e = cudaEventCreate(&c->evt0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->evt1);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&c->pinned_counts, 4096, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        tfbs_set_error("tfbs_ctx_create: %s", cudaGetErrorString(e));
        delete c;
        return TFBS_ERR_CUDA;
    }
    *out = c;
    return TFBS_OK;
}

int tfbs_ctx_destroy(tfbs_ctx *ctx)
{
    if (!ctx) return TFBS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_scratch(ctx->dense_fwd);
    free_scratch(ctx->dense_rev);
    free_scratch(ctx->hits);
    free_scratch(ctx->hit_count);
    free_scratch(ctx->cand);
    free_scratch(ctx->tracks);
    free_scratch(ctx->sites);
    free_scratch(ctx->misc);
    free_scratch(ctx->flush);
    free_scratch(ctx->thr);
    free_scratch(ctx->sort);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_mid) cudaEventDestroy(ctx->ev_mid);
    if (ctx->ev_direct) cudaEventDestroy(ctx->ev_direct);
    if (ctx->evt0) cudaEventDestroy(ctx->evt0);
    if (ctx->evt1) cudaEventDestroy(ctx->evt1);
    for (cudaEvent_t ev : ctx->pipe_events) cudaEventDestroy(ev);
    if (ctx->pinned_counts) cudaFreeHost(ctx->pinned_counts);
    if (ctx->stream_in) cudaStreamDestroy(ctx->stream_in);
    if (ctx->stream_out) cudaStreamDestroy(ctx->stream_out);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TFBS_OK;
}

int tfbs_ctx_sync(tfbs_ctx *ctx)
{
    TFBS_REQUIRE(ctx, "tfbs_ctx_sync: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return TFBS_OK;
}

int tfbs_last_kernel_ms(tfbs_ctx *ctx, float *ms)
{
    TFBS_REQUIRE(ctx && ms, "tfbs_last_kernel_ms: null argument");
    *ms = ctx->last_ms;
    return TFBS_OK;
}

int tfbs_last_kernel_launches(tfbs_ctx *ctx, int64_t *n)
{
    TFBS_REQUIRE(ctx && n, "tfbs_last_kernel_launches: null argument");
    *n = ctx->last_launches;
    return TFBS_OK;
}

int tfbs_total_kernel_launches(tfbs_ctx *ctx, int64_t *n)
{
    TFBS_REQUIRE(ctx && n, "tfbs_total_kernel_launches: null argument");
    *n = ctx->total_launches;
    return TFBS_OK;
}

int tfbs_last_scan_candidates(tfbs_ctx *ctx, int64_t *n)
{
    TFBS_REQUIRE(ctx && n, "tfbs_last_scan_candidates: null argument");
    *n = ctx->last_candidates;
    return TFBS_OK;
}

int tfbs_last_hits_ms(tfbs_ctx *ctx, float *scan_ms, float *verify_ms)
{
    TFBS_REQUIRE(ctx && scan_ms && verify_ms, "tfbs_last_hits_ms: null argument");
    *scan_ms = ctx->last_scan_ms;
    *verify_ms = ctx->last_verify_ms;
    return TFBS_OK;
}

int tfbs_last_direct_ms(tfbs_ctx *ctx, float *direct_ms)
{
    TFBS_REQUIRE(ctx && direct_ms, "tfbs_last_direct_ms: null argument");
    *direct_ms = ctx->last_direct_ms;
    return TFBS_OK;
}

int tfbs_timer_start(tfbs_ctx *ctx)
{
    TFBS_REQUIRE(ctx, "tfbs_timer_start: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    TFBS_CHECK_CUDA(cudaEventRecord(ctx->evt0, ctx->stream));
    return TFBS_OK;
}

int tfbs_timer_stop(tfbs_ctx *ctx, float *ms)
{
    TFBS_REQUIRE(ctx && ms, "tfbs_timer_stop: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    TFBS_CHECK_CUDA(cudaEventRecord(ctx->evt1, ctx->stream));
    TFBS_CHECK_CUDA(cudaEventSynchronize(ctx->evt1));
    TFBS_CHECK_CUDA(cudaEventElapsedTime(ms, ctx->evt0, ctx->evt1));
    return TFBS_OK;
}

int tfbs_host_alloc(void **ptr, int64_t bytes)
{
    TFBS_REQUIRE(ptr && bytes >= 0, "tfbs_host_alloc: bad argument");
    TFBS_CHECK_CUDA(cudaHostAlloc(ptr, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return TFBS_OK;
}

int tfbs_device_pci_bus_id(int device, char *buf, int32_t len)
{
    TFBS_REQUIRE(buf && len >= 16, "tfbs_device_pci_bus_id: buffer of at least 16 bytes needed");
    TFBS_CHECK_CUDA(cudaDeviceGetPCIBusId(buf, len, device));
    return TFBS_OK;
}

int tfbs_host_register(void *ptr, int64_t bytes)
{
    TFBS_REQUIRE(ptr && bytes > 0, "tfbs_host_register: bad argument");
    TFBS_CHECK_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return TFBS_OK;
}

int tfbs_host_unregister(void *ptr)
{
    TFBS_REQUIRE(ptr, "tfbs_host_unregister: null argument");
    TFBS_CHECK_CUDA(cudaHostUnregister(ptr));
    return TFBS_OK;
}

int tfbs_host_free(void *ptr)
{
    if (ptr) TFBS_CHECK_CUDA(cudaFreeHost(ptr));
    return TFBS_OK;
}

int tfbs_flush_l2(tfbs_ctx *ctx)
{
    TFBS_REQUIRE(ctx, "tfbs_flush_l2: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    int rc = tfbs_scratch_reserve(ctx->flush, bytes);
    if (rc) return rc;
    static uint32_t tick = 0;
    flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4 *)ctx->flush.ptr, bytes / 16, ++tick);
    TFBS_CHECK_CUDA(cudaGetLastError());
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return TFBS_OK;
}

namespace {
// Shared-memory load bandwidth probe: pointer chasing through a 32 KB table of 256 rows x 32 words.
// Every word of a row holds the byte offset of the next row, so the 32 lanes of a warp always read the
// 32 consecutive words of one row (conflict-free, one 128 B wavefront per LDS.32 — the access pattern of
// the hits kernel's table look-ups) and a load costs one LDS + one IADD.  8 independent chains per
// thread x 32 warps keep far more loads in flight than latency x bandwidth needs.
__global__ void __launch_bounds__(1024, 1) smem_probe_kernel(int iters, unsigned int *sink)
{
    __shared__ uint32_t tab[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = (uint32_t)((((i >> 5) * 37 + 11) & 255) * 128);
    __syncthreads();
    const uint32_t lane_base = (uint32_t)__cvta_generic_to_shared(tab) + (threadIdx.x & 31u) * 4u;
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = (uint32_t)(((threadIdx.x >> 5) * 8 + k) & 255) * 128u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a[k]) : "r"(a[k] + lane_base));
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) x ^= a[k];
    if (x == (uint32_t)iters) *sink = x;  // (x is a multiple of 128, iters is not) keeps the chains alive
}
}  // namespace

int tfbs_measure_smem_gbs(tfbs_ctx *ctx, float *gbs)
{
    TFBS_REQUIRE(ctx && gbs, "tfbs_measure_smem_gbs: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    int rc = tfbs_scratch_reserve(ctx->misc, 64);
    if (rc) return rc;
    const int iters = 20000;
    float best = 0.f;
    for (int rep = 0; rep < 4; rep++) {
        tfbs_begin(ctx);
        smem_probe_kernel<<<ctx->sm_count, 1024, 0, ctx->stream>>>(iters, (unsigned int *)ctx->misc.ptr);
        tfbs_launched(ctx);
        rc = tfbs_end(ctx);
        if (rc) return rc;
        const double bytes = (double)ctx->sm_count * 1024.0 * 8.0 * iters * 4.0;
        const float g = (float)(bytes / (ctx->last_ms * 1e-3) / 1e9);
        if (rep > 0 && g > best) best = g;  // first launch is the warm-up
    }
    *gbs = best;
    return TFBS_OK;
}

int tfbs_device_free(tfbs_ctx *ctx, void *dev_ptr)
{
    TFBS_REQUIRE(ctx, "tfbs_device_free: null argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    if (dev_ptr) TFBS_CHECK_CUDA(cudaFree(dev_ptr));
    return TFBS_OK;
}

int tfbs_device_download(tfbs_ctx *ctx, const void *dev_ptr, void *host, int64_t bytes)
{
    TFBS_REQUIRE(ctx && (bytes == 0 || (dev_ptr && host)) && bytes >= 0, "tfbs_device_download: bad argument");
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    TFBS_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bytes > 0) TFBS_CHECK_CUDA(cudaMemcpy(host, dev_ptr, (size_t)bytes, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}

}  // extern "C"
