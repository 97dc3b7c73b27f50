// lucifit_fit_kernel.cuh -- the fused per-spaxel kernel (one warp per spaxel) and its launcher.
#pragma once
#include <cuda_runtime.h>
#include "lucifit_core.cuh"

#define LF_WARPS_PER_BLOCK 8

struct LfLaunch {
    LfCfg cfg;
    long long n;
    const float* cube;
    const long long* index;
    const float* theta;
    const float* vel0;
    const float* broad0;
    float* out_rows;
    float* out_sol;
    float* out_unc;
    int* status;
    int off_u, off_v, per_warp, cont_cap, xo_bytes;   // shared-memory layout (bytes)
    cudaStream_t stream;
    int sm_count;
};
typedef cudaError_t (*lf_launch_fn)(const LfLaunch&);

template <int MODEL, int L, int NG>
__global__ void __launch_bounds__(LF_WARPS_PER_BLOCK * 32, 2)
lf_fit_kernel(const __grid_constant__ LfLaunch a)
{
    extern __shared__ __align__(16) unsigned char lf_smem[];
    float* xo_s = reinterpret_cast<float*>(lf_smem);
    const int nfit = a.cfg.fit_hi - a.cfg.fit_lo;
    for (int i = threadIdx.x; i < nfit; i += blockDim.x) xo_s[i] = a.cfg.xo[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    unsigned char* base = lf_smem + a.xo_bytes + (size_t)warp * a.per_warp;
    float* spec = reinterpret_cast<float*>(base);
    LfScratch sc;
    sc.dvals = reinterpret_cast<double*>(base + a.off_u);
    sc.ddev = sc.dvals + a.cont_cap;
    sc.alive = reinterpret_cast<unsigned char*>(sc.ddev + a.cont_cap);
    sc.hess = reinterpret_cast<double*>(base + a.off_u);
    sc.st = reinterpret_cast<float*>(base + a.off_v);
    sc.tile = sc.st;
    LfWarp<32> w;
    constexpr int K = 7 * L + 5, PF = 3 * L + 1;
    const long long stride = (long long)gridDim.x * LF_WARPS_PER_BLOCK;
    for (long long s = (long long)blockIdx.x * LF_WARPS_PER_BLOCK + warp; s < a.n; s += stride) {
        long long row = a.index ? a.index[s] : s;
        lf_fit_spaxel<MODEL, L, NG, NG, 32>(a.cfg, w, xo_s, a.cube + row * (long long)a.cfg.nz, a.theta[s],
                                            a.vel0 ? a.vel0[s] : 0.0f, a.broad0 ? a.broad0[s] : 0.0f, spec, sc,
                                            a.out_rows + s * K, a.status + s, a.out_sol ? a.out_sol + s * PF : nullptr,
                                            a.out_unc ? a.out_unc + s * PF : nullptr);
        __syncwarp();
    }
}

template <int MODEL, int L, int NG>
cudaError_t lf_launch(const LfLaunch& a)
{
    auto kern = lf_fit_kernel<MODEL, L, NG>;
    size_t smem = (size_t)a.xo_bytes + (size_t)LF_WARPS_PER_BLOCK * a.per_warp;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LF_WARPS_PER_BLOCK * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    long long want = (a.n + LF_WARPS_PER_BLOCK - 1) / LF_WARPS_PER_BLOCK;
    long long cap = (long long)a.sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, LF_WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    return cudaGetLastError();
}
