// lucifit_fit_kernel.cuh -- the four stage kernels of the fit path (one warp per spaxel, grid-stride over the
// slice) and their launcher.  Each kernel keeps one small hot loop resident in the SM's instruction cache; a spaxel
// travels between them as a state record in the caller's workspace (lucifit_core.cuh, LF_SO_*).
#pragma once
#include <nvtx3/nvToolsExt.h>   // header-only; the ranges cost nothing unless a profiler is attached
#include <cuda_runtime.h>
#include "lucifit_core.cuh"

// Warps per CTA.  Measured on the benchmark cube (B200, 1 M spaxels): consecutive spaxels go to consecutive warps of a
// CTA, and larger CTAs keep neighbouring spaxels (neighbouring DRAM pages) on one SM -- the LM kernel runs 57.1 / 54.7 /
// 52.5 ms with 4 / 8 / 16 warps per CTA at the same 16 resident warps per SM, the prep kernel 3.71 / 3.48 / 3.42 ms with
// 8 / 16 / 32, the output kernel 6.42 / 5.84 / 5.40 ms (with the 80 registers it used to get, 8 was its best size).
#ifndef LF_PREP_WARPS
#define LF_PREP_WARPS 32
#endif
#ifndef LF_PREP_LOCKSTEP
#define LF_PREP_LOCKSTEP 1
#endif
// The LM kernel keeps ONE resident CTA per SM and sizes it by the normal equations: up to 8 unknowns (44 accumulators)
// fit the 128 registers that 16 warps leave per thread; larger systems get 12 or 8 warps and the registers they need.
#ifndef LF_LM_WARPS_SMALL
#define LF_LM_WARPS_SMALL 16
#endif
#ifndef LF_LM_MINB_SMALL
#define LF_LM_MINB_SMALL 1
#endif
template <int NA> struct LfLmBlocks {
    enum { W = NA <= 44 ? LF_LM_WARPS_SMALL : NA <= 77 ? 12 : 8, V = NA <= 44 ? LF_LM_MINB_SMALL : 1 };
};
// The uncertainty kernel is bound by its per-warp shared memory (13 L variant records, the variant / sum tiles): one CTA per
// SM with as many warps as fit 227 KB -- 20 up to 5 lines (11.2 KB each at L = 5), 16 at 6, 12 beyond.  Round 1 measured
// 96.2 / 91.9 / 89.6 / 91.3 ms per 2^20 spaxels with 4 / 8 / 16 / 18 warps per CTA; with the leaner code of round 2, 20 warps
// at 96 registers (76 bytes of spills) beat 16 at 128: 31.6 against 34.4 ms per 2^19 spaxels.  The launcher steps down to
// 16 / 12 / 8 warps when the scratch of a configuration does not fit (lf_unc_warps_for).
#ifndef LF_UNC_WARPS
#define LF_UNC_WARPS 20
#endif
template <int L> struct LfUncWarps { enum { W = L <= 5 ? LF_UNC_WARPS : (L <= 6 ? (LF_UNC_WARPS < 16 ? LF_UNC_WARPS : 16) : (LF_UNC_WARPS < 12 ? LF_UNC_WARPS : 12)) }; };
#ifndef LF_OUT_WARPS
#define LF_OUT_WARPS 32
#endif
// The prep and output kernels take their CTA size at launch (lf_pick_warps): consecutive spaxels go to consecutive warps, so
// a larger CTA keeps neighbouring spectra (neighbouring DRAM pages, one copy of the xo table) on one SM -- the output kernel
// takes 6.42 / 5.84 / 5.40 ms per 2^20 spaxels with 8 / 16 / 32 warps per CTA at the same 64-register budget.  LF_*_WARPS
// is the largest size tried; long spectra fall back to 16 or 8 warps to stay inside the shared memory of an SM.
// (64 registers either way: 32 resident warps per SM)
#define LF_PREP_BOUNDS __launch_bounds__(LF_PREP_WARPS * 32, 32 / LF_PREP_WARPS)
#define LF_OUT_BOUNDS __launch_bounds__(LF_OUT_WARPS * 32, 32 / LF_OUT_WARPS)

struct LfLaunch {
    LfCfg cfg;
    LfLayout ly;
    long long n;                 // spaxels of this slice
    const float* cube;
    const long long* index;      // already offset to the slice, or null
    long long index_base;        // row of the slice's first spaxel when index is null
    const float* theta;
    const float* vel0;
    const float* broad0;
    float* out_rows;
    float* out_sol;
    float* out_unc;
    int* status;
    float* state;                // [n][cfg.state_floats]
    double* unc_ws;              // [n][3L+1] or null
    int xo_bytes;
    cudaStream_t stream;
    int sm_count;
    int stage_rows;              // out_rows is peer-mapped memory of another GPU: rows leave as whole coalesced stores
};
typedef cudaError_t (*lf_launch_fn)(const LfLaunch&, int stages);

__device__ __forceinline__ const float* lf_stage_xo(const LfLaunch& a, unsigned char* smem)
{
    float* xo_s = reinterpret_cast<float*>(smem);
    const int nfit = a.cfg.fit_hi - a.cfg.fit_lo;
    for (int i = threadIdx.x; i < nfit; i += blockDim.x) xo_s[i] = a.cfg.xo[i];
    __syncthreads();
    return xo_s;
}

template <int MODEL, int L, int NG>
__global__ void LF_PREP_BOUNDS
lf_prep_kernel(const __grid_constant__ LfLaunch a)
{
    extern __shared__ __align__(16) unsigned char lf_smem[];
    const int warp = threadIdx.x >> 5;
    LfWork wk = lf_work_at(LF_STAGE_PREP, a.ly, lf_smem + (size_t)warp * a.ly.prep_bytes);
    LfWarp<32> w;
    const int nwarps = blockDim.x >> 5;
    const long long stride = (long long)gridDim.x * nwarps;
    // The stage is ~60 KB of straight-line code walked once per spaxel, with near-identical work for every spaxel: the
    // warps of a CTA start each spaxel together (one CTA barrier per round, uniform trip count) so that they fetch the
    // same instructions at about the same time instead of thrashing the instruction cache (ncu: 1.9 `no_instruction`
    // stall cycles per issued instruction without the barrier).
    for (long long s0 = (long long)blockIdx.x * nwarps; s0 < a.n; s0 += stride) {
#if LF_PREP_LOCKSTEP
        __syncthreads();
#endif
        const long long s = s0 + warp;
        if (s >= a.n) continue;
        const long long row = a.index ? a.index[s] : a.index_base + s;
        const int pnv = a.cfg.prior_groups ? a.cfg.nv : 1, pns = a.cfg.prior_groups ? a.cfg.ns : 1;
        lf_stage_prep<MODEL, L, NG, NG, 32>(a.cfg, w, a.cube + row * (long long)a.cfg.nz, a.theta[s],
                                            a.vel0 ? a.vel0 + s * pnv : nullptr, a.broad0 ? a.broad0 + s * pns : nullptr,
                                            wk, a.state + s * (long long)a.cfg.state_floats);
    }
}

template <int MODEL, int L, int NG, int PHASE>
__global__ void __launch_bounds__((LfLmBlocks<LfDims<MODEL, L, NG, NG>::NA>::W) * 32, (LfLmBlocks<LfDims<MODEL, L, NG, NG>::NA>::V))
lf_lm_kernel(const __grid_constant__ LfLaunch a)
{
    extern __shared__ __align__(16) unsigned char lf_smem[];
    constexpr int WARPS = LfLmBlocks<LfDims<MODEL, L, NG, NG>::NA>::W;
    const float* xo_s = lf_stage_xo(a, lf_smem);
    const int warp = threadIdx.x >> 5;
    LfWork wk = lf_work_at(LF_STAGE_LM, a.ly, lf_smem + a.xo_bytes + (size_t)warp * a.ly.lm_bytes);
    LfWarp<32> w;
    const long long stride = (long long)gridDim.x * WARPS;
    for (long long s = (long long)blockIdx.x * WARPS + warp; s < a.n; s += stride) {
        const long long row = a.index ? a.index[s] : a.index_base + s;
        if (PHASE == 2) {      // only the spaxels the first kernel left unfinished (their rows are not even addressed otherwise)
            const int st = reinterpret_cast<const int*>(a.state + s * (long long)a.cfg.state_floats)[LF_SO_STATUS];
            if ((st & (LF_ST_MAXITER | LF_ST_CONVERGED | LF_ST_BADMASK)) != LF_ST_MAXITER) continue;
        }
        lf_stage_lm<MODEL, L, NG, NG, 32, PHASE>(a.cfg, w, xo_s, a.cube + row * (long long)a.cfg.nz, wk,
                                                 a.state + s * (long long)a.cfg.state_floats);
    }
}

template <int MODEL, int L, int NG>
__global__ void __launch_bounds__(LfUncWarps<L>::W * 32, 1)
lf_unc_kernel(const __grid_constant__ LfLaunch a)
{
    extern __shared__ __align__(16) unsigned char lf_smem[];
    const int WARPS = blockDim.x >> 5;
    const float* xo_s = lf_stage_xo(a, lf_smem);
    const int warp = threadIdx.x >> 5;
    LfWork wk = lf_work_at(LF_STAGE_UNC, a.ly, lf_smem + a.xo_bytes + (size_t)warp * a.ly.un_bytes);
    LfWarp<32> w;
    const long long stride = (long long)gridDim.x * WARPS;
    for (long long s = (long long)blockIdx.x * WARPS + warp; s < a.n; s += stride) {
        const long long row = a.index ? a.index[s] : a.index_base + s;
        lf_stage_unc<MODEL, L, NG, NG, 32>(a.cfg, w, xo_s, a.cube + row * (long long)a.cfg.nz, wk,
                                           a.state + s * (long long)a.cfg.state_floats, a.unc_ws + s * (3 * L + 1));
    }
}

template <int MODEL, int L, int NG, bool STAGE_ROW>
__global__ void LF_OUT_BOUNDS
lf_out_kernel(const __grid_constant__ LfLaunch a)
{
    extern __shared__ __align__(16) unsigned char lf_smem[];
    const float* xo_s = lf_stage_xo(a, lf_smem);
    const int warp = threadIdx.x >> 5;
    LfWork wk = lf_work_at(LF_STAGE_OUT, a.ly, lf_smem + a.xo_bytes + (size_t)warp * a.ly.out_bytes);
    LfWarp<32> w;
    constexpr int K = 7 * L + 5, PF = 3 * L + 1;
    const int nwarps = blockDim.x >> 5;
    const long long stride = (long long)gridDim.x * nwarps;
    for (long long s = (long long)blockIdx.x * nwarps + warp; s < a.n; s += stride) {
        const long long row = a.index ? a.index[s] : a.index_base + s;
        lf_stage_out<MODEL, L, NG, NG, 32, STAGE_ROW>(a.cfg, w, xo_s, a.cube + row * (long long)a.cfg.nz, wk,
                                           a.state + s * (long long)a.cfg.state_floats,
                                           a.unc_ws ? a.unc_ws + s * PF : nullptr, a.out_rows + s * K, a.status + s,
                                           a.out_sol ? a.out_sol + s * PF : nullptr,
                                           a.out_unc ? a.out_unc + s * PF : nullptr);
    }
}

template <typename K>
static cudaError_t lf_launch_one(K kern, const LfLaunch& a, int warps, size_t smem)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    long long want = (a.n + warps - 1) / warps;
    long long cap = (long long)a.sm_count * occ;          // one resident wave, grid-stride inside
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, warps * 32, smem, a.stream>>>(a);
    return cudaGetLastError();
}

// Largest CTA (in warps) of {max_warps, 16, 8} whose dynamic shared memory stays within LF_SMEM_BUDGET.
#define LF_SMEM_BUDGET (200 * 1024)
static inline int lf_pick_warps(int max_warps, size_t fixed_bytes, size_t bytes_per_warp)
{
    int w = max_warps;
    while (w > 8 && fixed_bytes + (size_t)w * bytes_per_warp > (size_t)LF_SMEM_BUDGET) w >>= 1;
    return w;
}

// The uncertainty kernel's CTA: max_warps, else 16 / 12 / 8, whichever first fits the 227 KB of dynamic shared memory.
static inline int lf_unc_warps_for(int max_warps, size_t fixed_bytes, size_t bytes_per_warp)
{
    const int cand[4] = {max_warps, 16, 12, 8};
    for (int i = 0; i < 4; ++i)
        if (cand[i] <= max_warps && fixed_bytes + (size_t)cand[i] * bytes_per_warp <= (size_t)227 * 1024) return cand[i];
    return 8;
}

// Enqueues the requested stages (bit 0 prep, 1 lm, 2 unc, 3 out) for one slice, in pipeline order.
template <int MODEL, int L, int NG>
cudaError_t lf_launch(const LfLaunch& a, int stages)
{
    cudaError_t e = cudaSuccess;
    struct Range { explicit Range(const char* n) { nvtxRangePushA(n); } ~Range() { nvtxRangePop(); } };
    if (stages & 1) {
        Range r("lucifit:prep");
        const int pw = lf_pick_warps(LF_PREP_WARPS, 0, (size_t)a.ly.prep_bytes);
        e = lf_launch_one(lf_prep_kernel<MODEL, L, NG>, a, pw, (size_t)pw * a.ly.prep_bytes);
        if (e != cudaSuccess) return e;
    }
    if (stages & 2) {
        Range r("lucifit:lm");
        constexpr int LMW = LfLmBlocks<LfDims<MODEL, L, NG, NG>::NA>::W;
        e = lf_launch_one(lf_lm_kernel<MODEL, L, NG, 1>, a, LMW, (size_t)a.xo_bytes + (size_t)LMW * a.ly.lm_bytes);
        if (e != cudaSuccess) return e;
        if (!a.cfg.freeze) {
            e = lf_launch_one(lf_lm_kernel<MODEL, L, NG, 2>, a, LMW, (size_t)a.xo_bytes + (size_t)LMW * a.ly.lm_bytes);
            if (e != cudaSuccess) return e;
        }
    }
    if ((stages & 4) && a.unc_ws) {
        Range r("lucifit:unc");
        const int uw = lf_unc_warps_for(LfUncWarps<L>::W, (size_t)a.xo_bytes, (size_t)a.ly.un_bytes);
        e = lf_launch_one(lf_unc_kernel<MODEL, L, NG>, a, uw, (size_t)a.xo_bytes + (size_t)uw * a.ly.un_bytes);
        if (e != cudaSuccess) return e;
    }
    if (stages & 8) {
        Range r("lucifit:out");
        const int ow = lf_pick_warps(LF_OUT_WARPS, (size_t)a.xo_bytes, (size_t)a.ly.out_bytes);
        const size_t smem = (size_t)a.xo_bytes + (size_t)ow * a.ly.out_bytes;
        e = a.stage_rows ? lf_launch_one(lf_out_kernel<MODEL, L, NG, true>, a, ow, smem)
                         : lf_launch_one(lf_out_kernel<MODEL, L, NG, false>, a, ow, smem);
    }
    return e;
}
