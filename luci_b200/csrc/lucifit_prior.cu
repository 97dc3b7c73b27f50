// lucifit_prior.cu -- kernel of lucifit_estimate_prior (one warp per spaxel, grid-stride; lucifit_prior.cuh).
#include <cuda_runtime.h>
#include "lucifit_prior.cuh"

#define LF_PRIOR_WARPS 8

__global__ void __launch_bounds__(LF_PRIOR_WARPS * 32)
lf_prior_kernel(const __grid_constant__ LfPriorCfg c, long long n, const float* __restrict__ cube,
                const long long* __restrict__ index, const float* __restrict__ theta, float* vel, float* broad)
{
    extern __shared__ __align__(16) float lf_prior_smem[];
    const int warp = threadIdx.x >> 5;
    const int per_warp = 2 * c.nfit + c.nv;
    float* z = lf_prior_smem + (size_t)warp * per_warp;
    float* sm = z + c.nfit;
    float* score = sm + c.nfit;
    LfWarp<32> w;
    const long long stride = (long long)gridDim.x * LF_PRIOR_WARPS;
    for (long long s = (long long)blockIdx.x * LF_PRIOR_WARPS + warp; s < n; s += stride) {
        const long long row = index ? index[s] : s;
        lf_estimate_prior<32>(c, w, cube + row * (long long)c.nz, theta[s], z, sm, score, vel + s, broad + s);
    }
}

extern "C" cudaError_t lf_run_prior(const LfPriorCfg* c, long long n, const float* cube, const long long* index,
                                    const float* theta, float* vel, float* broad, int sm_count, cudaStream_t stream)
{
    const size_t smem = sizeof(float) * (size_t)LF_PRIOR_WARPS * (2 * c->nfit + c->nv);
    cudaError_t e = cudaFuncSetAttribute(lf_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_prior_kernel, LF_PRIOR_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    long long want = (n + LF_PRIOR_WARPS - 1) / LF_PRIOR_WARPS, cap = (long long)sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    lf_prior_kernel<<<grid, LF_PRIOR_WARPS * 32, smem, stream>>>(*c, n, cube, index, theta, vel, broad);
    return cudaGetLastError();
}
