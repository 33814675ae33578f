// tmem_probe.cu — is Tensor Memory usable as a per-lane look-up table (dynamic column, one value per
// lane), and how fast is tcgen05.ld next to LDS?   Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t ldtm(uint32_t taddr)
{
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    return v;
}
__device__ __forceinline__ void ldtm_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int kCols = 256;

// mode 0: 8 LDTM / iteration; mode 1: 8 LDS; mode 2: 4 LDTM + 8 LDS interleaved
__global__ void __launch_bounds__(256) probe(uint32_t *out, long long *cycles, int iters, int mode)
{
    __shared__ uint32_t s_taddr;
    __shared__ uint32_t s_tab[kCols * 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kCols * 32; i += blockDim.x) s_tab[i] = (uint32_t)i * 2654435761u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_taddr)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = s_taddr;
    const uint32_t my = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
        for (int c = 0; c < kCols; c++) {
            const uint32_t val = s_tab[c * 32 + lane] ^ (uint32_t)(warp << 28);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(my + c), "r"(val));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // correctness: every warp reads all columns of its quadrant
    uint32_t bad = 0;
    for (int c = 0; c < kCols; c++) {
        const uint32_t v = ldtm(my + c);
        ldtm_wait();
        bad += v != (s_tab[c * 32 + lane] ^ (uint32_t)((warp & 3) << 28));
    }
    // throughput
    uint32_t acc = 0, x = 12345u + warp * 77u + blockIdx.x;
    const uint32_t lane_addr = smem_u32(s_tab) + lane * 4;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        // cheap, warp-uniform, iteration-dependent columns (no long dependency chain)
        const uint32_t b = (uint32_t)it * 7u + warp * 13u;
        if (mode == 0) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = ldtm(my + ((b + k * 37u) & (kCols - 1)));
            ldtm_wait();
#pragma unroll
            for (int k = 0; k < 8; k++) acc += v[k];
        } else if (mode == 1) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(lane_addr + ((b + k * 37u) & (kCols - 1)) * 128));
#pragma unroll
            for (int k = 0; k < 8; k++) acc += v[k];
        } else {
            uint32_t v[8], t[4];
#pragma unroll
            for (int k = 0; k < 4; k++) t[k] = ldtm(my + ((b + k * 37u) & (kCols - 1)));
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(lane_addr + ((b + k * 41u + 3u) & (kCols - 1)) * 128));
            ldtm_wait();
#pragma unroll
            for (int k = 0; k < 8; k++) acc += v[k];
#pragma unroll
            for (int k = 0; k < 4; k++) acc += t[k];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kCols));
    out[blockIdx.x * blockDim.x + tid] = acc + (bad << 24);
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (bad) atomicAdd(&out[gridDim.x * blockDim.x], bad);
}

int main()
{
    const int blocks = 296, threads = 256, iters = 20000;
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, (blocks * threads + 1) * 4);
    cudaMalloc(&cyc, blocks * 8);
    for (int mode = 0; mode < 3; mode++) {
        cudaMemset(out, 0, (blocks * threads + 1) * 4);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<<<blocks, threads>>>(out, cyc, iters, mode);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        uint32_t bad = 0;
        long long c0 = 0;
        cudaMemcpy(&bad, out + blocks * threads, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        // per SM: 2 CTAs x 8 warps; lookups per warp-iteration: mode0 4 TMEM, mode1 4 LDS, mode2 2 TMEM + 4 LDS
        const double per_sm_warp_lookups = 2.0 * 8 * iters * (mode == 2 ? 12 : 8);
        printf("mode %d: %s, mismatches %u, %.3f ms, CTA0 cycles %lld, warp-lookups/clk/SM = %.3f\n", mode,
               cudaGetErrorString(err), bad, ms, c0, per_sm_warp_lookups / (double)c0);
    }
    return 0;
}
