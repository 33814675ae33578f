// classify.cu — scaler + SVC decision function over a feature matrix.
//
// Replaces the inference arithmetic of predict.py:286-319: the MinMaxScaler -> StandardScaler
// pipeline fitted in cross_validate_model.py:182-189 applied to the prediction features, then
// sklearn SVC.decision_function (libsvm: sum_i dual_i K(sv_i, x) + intercept; kernels "linear" and
// "rbf" are the two the reference's grid searches, cross_validate_model.py:204-208) and, for the
// linear kernel, the distance from the hyperplane decision / ||w|| (predict.py:314-319).  All in
// double, so results agree with sklearn to ~1e-12; the feature matrix never has to leave the GPU.
//
// One warp per sample: the lanes stride over the features, the transformed sample is staged once in
// shared memory, every support vector costs F / 32 fused multiply-adds per lane and one warp reduction.
#include <cmath>

#include "common.cuh"

namespace {

constexpr int kSvcThreads = 256;

struct SvcParams {
    const double *X;         // [S][F] raw features
    int64_t S;
    int32_t F;
    const double *mm_scale;  // MinMaxScaler.scale_, .min_       : x1 = x * scale + min
    const double *mm_min;
    const double *st_mean;   // StandardScaler.mean_, .scale_     : x2 = (x1 - mean) / scale
    const double *st_scale;
    const double *sv;        // [nSV][F] support vectors (rbf) or coef_ as one row (linear)
    const double *dual;      // [nSV] dual coefficients (rbf)
    int32_t n_sv;
    int32_t kernel;          // 0 linear, 1 rbf
    double gamma, intercept;
    double *out;             // [S]
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

__global__ void __launch_bounds__(kSvcThreads) svc_decision_kernel(const SvcParams P)
{
    extern __shared__ double s_x[];  // [warps][F]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *x = s_x + (size_t)warp * P.F;
    const int64_t n_warps = ((int64_t)gridDim.x * kSvcThreads) >> 5;
    for (int64_t s = (((int64_t)blockIdx.x * kSvcThreads) >> 5) + warp; s < P.S; s += n_warps) {
        const double *row = P.X + (size_t)s * P.F;
        for (int f = lane; f < P.F; f += 32) {
            const double a = row[f] * __ldg(P.mm_scale + f) + __ldg(P.mm_min + f);
            x[f] = (a - __ldg(P.st_mean + f)) / __ldg(P.st_scale + f);
        }
        __syncwarp();
        double acc = 0.0;
        if (P.kernel == 0) {
            double d = 0.0;
            for (int f = lane; f < P.F; f += 32) d += x[f] * __ldg(P.sv + f);
            acc = warp_sum(d);
        } else {
            for (int i = 0; i < P.n_sv; i++) {
                const double *v = P.sv + (size_t)i * P.F;
                double d = 0.0;
                for (int f = lane; f < P.F; f += 32) {
                    const double t = x[f] - __ldg(v + f);
                    d += t * t;
                }
                d = warp_sum(d);
                acc += __ldg(P.dual + i) * exp(-P.gamma * d);
            }
        }
        if (lane == 0) P.out[s] = acc + P.intercept;
        __syncwarp();
    }
}

}  // namespace

extern "C" int tfbs_svc_decision(tfbs_ctx *ctx, const double *X_host, int64_t S, int32_t F, const double *mm_scale,
                                 const double *mm_min, const double *st_mean, const double *st_scale, int32_t kernel,
                                 const double *sv, const double *dual, int32_t n_sv, double gamma, double intercept,
                                 double *out_host)
{
    TFBS_REQUIRE(ctx && S >= 0 && F >= 1 && mm_scale && mm_min && st_mean && st_scale && sv && (S == 0 || (X_host && out_host)),
                 "tfbs_svc_decision: null argument");
    TFBS_REQUIRE(kernel == 0 || (kernel == 1 && dual && n_sv >= 1), "tfbs_svc_decision: kernel must be 0 (linear) or 1 (rbf)");
    TFBS_REQUIRE((size_t)F * 8 * (kSvcThreads / 32) <= 200 * 1024, "tfbs_svc_decision: %d features do not fit shared memory", F);
    if (S == 0) return TFBS_OK;
    TFBS_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int rows = kernel == 0 ? 1 : n_sv;
    const size_t nX = (size_t)S * F, nP = (size_t)4 * F, nSV = (size_t)rows * F, nD = kernel ? (size_t)n_sv : 0;
    int rc = tfbs_scratch_reserve(ctx->sites, (nX + nP + nSV + nD + (size_t)S) * 8);
    if (rc) return rc;
    double *d = (double *)ctx->sites.ptr;
    SvcParams P;
    P.X = d;
    P.mm_scale = d + nX;
    P.mm_min = P.mm_scale + F;
    P.st_mean = P.mm_min + F;
    P.st_scale = P.st_mean + F;
    P.sv = P.st_scale + F;
    P.dual = P.sv + nSV;
    P.out = (double *)(P.dual + nD);
    P.S = S;
    P.F = F;
    P.n_sv = n_sv;
    P.kernel = kernel;
    P.gamma = gamma;
    P.intercept = intercept;
    cudaStream_t st = ctx->stream;
    TFBS_CHECK_CUDA(cudaMemcpyAsync((void *)P.X, X_host, nX * 8, cudaMemcpyHostToDevice, st));
    TFBS_CHECK_CUDA(cudaMemcpyAsync((void *)P.mm_scale, mm_scale, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    TFBS_CHECK_CUDA(cudaMemcpyAsync((void *)P.mm_min, mm_min, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    TFBS_CHECK_CUDA(cudaMemcpyAsync((void *)P.st_mean, st_mean, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    TFBS_CHECK_CUDA(cudaMemcpyAsync((void *)P.st_scale, st_scale, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    TFBS_CHECK_CUDA(cudaMemcpyAsync((void *)P.sv, sv, nSV * 8, cudaMemcpyHostToDevice, st));
    if (nD) TFBS_CHECK_CUDA(cudaMemcpyAsync((void *)P.dual, dual, nD * 8, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)F * 8 * (kSvcThreads / 32);
    TFBS_CHECK_CUDA(cudaFuncSetAttribute(svc_decision_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (S + kSvcThreads / 32 - 1) / (kSvcThreads / 32);
    if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
    tfbs_begin(ctx);
    svc_decision_kernel<<<(unsigned)blocks, kSvcThreads, smem, st>>>(P);
    tfbs_launched(ctx);
    rc = tfbs_end(ctx);
    if (rc) return rc;
    TFBS_CHECK_CUDA(cudaMemcpy(out_host, P.out, (size_t)S * 8, cudaMemcpyDeviceToHost));
    return TFBS_OK;
}
