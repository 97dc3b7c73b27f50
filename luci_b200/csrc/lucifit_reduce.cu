// lucifit_reduce.cu -- HBM-bound kernels of the path: spectral-axis reductions (deep image, SNR map), spatial
// binning and the synthetic-cube generator.
//
//   deep image   LuciBase.py:171-178        nansum over z of every spaxel
//   SNR map      LuciBase.py:1142-1172      per-spaxel max / mean / mean|.| / windowed sum / windowed std
//                                           (+ iterative 1-sigma clip minimum for method 2)
//   bin cube     LUCI/LuciUtility.py:332-341  b x b nansum of spectra
//   sim cube     LUCI/LuciSim.py:159-203    LuciSim.Spectrum.create_spectrum for every spaxel
//
// Layout: cube[spaxel][z], z contiguous (cube_final[x, y, z], LuciBase.py:122).  One warp streams one spaxel at a
// time with 16-byte loads where the row alignment allows it; a block stages nothing in shared memory for the
// plain reductions (each byte is used once), the sigma-clip of method 2 keeps the spectrum in shared memory.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "lucifit_core.cuh"

#define RED_WARPS 8
#define RED_UNROLL 8

struct SnrAcc {                   // four independent chains (one per float4 component) keep the FP32 adds pipelined
    float4 sum, asum;             // nansum, nansum|.|
    float mx;                     // nanmax
    int nan;                      // NaN count
};

__device__ __forceinline__ void snr_take1(float& sum, float& asum, float& mx, int& nan, float v)
{
    const bool bad = v != v;
    const float z = bad ? 0.0f : v;
    sum += z; asum += fabsf(z);
    mx = fmaxf(mx, bad ? -INFINITY : v);
    nan += bad;
}
__device__ __forceinline__ void snr_take4(SnrAcc& a, const float4& q)
{
    int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    snr_take1(a.sum.x, a.asum.x, m0, n0, q.x); snr_take1(a.sum.y, a.asum.y, m1, n1, q.y);
    snr_take1(a.sum.z, a.asum.z, m2, n2, q.z); snr_take1(a.sum.w, a.asum.w, m3, n3, q.w);
    a.mx = fmaxf(fmaxf(a.mx, fmaxf(m0, m1)), fmaxf(m2, m3));
    a.nan += (n0 + n1) + (n2 + n3);
}

// MODE 0: deep only; MODE 1: SNR method 1 (+deep).  The streaming loop carries only the whole-spectrum moments;
// the noise-window variance is accumulated in float64 from samples loaded up front (independent of the streaming
// loads, so nothing serialises behind them), which keeps the per-sample window tests out of the loop that has to
// keep up with HBM.
template <int MODE>
__global__ void __launch_bounds__(RED_WARPS * 32)
lf_reduce_kernel(const float* __restrict__ cube, long long n_spaxel, int nz, long long n_live, int flo, int fhi,
                 int nlo, int nhi, float* __restrict__ snr, float* __restrict__ deep)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = (long long)gridDim.x * RED_WARPS;
    for (long long s = (long long)blockIdx.x * RED_WARPS + warp; s < n_spaxel; s += stride) {
        const float* row = cube + s * (long long)nz;
        SnrAcc a;
        a.sum = a.asum = make_float4(0.f, 0.f, 0.f, 0.f); a.mx = -INFINITY; a.nan = 0;
        double n1 = 0.0, n2 = 0.0;
        int ncnt = 0;
        if (MODE == 1) {
            for (int z = nlo + lane; z < nhi; z += 32) {
                const float v = __ldg(row + z);
                if (v == v) { n1 += (double)v; n2 += (double)v * (double)v; ncnt += 1; }
            }
        }
        // peel to 16-byte alignment, then float4 body
        int head = (int)(((16 - ((uintptr_t)row & 15)) & 15) >> 2);
        if (head > nz) head = nz;
        if (lane < head) {
            float v = row[lane];
            if (MODE == 0) { if (v == v) a.sum.x += v; } else snr_take1(a.sum.x, a.asum.x, a.mx, a.nan, v);
        }
        const int nvec = (nz - head) >> 2;
        const float4* r4 = reinterpret_cast<const float4*>(row + head);
        // RED_UNROLL independent 16-byte loads per lane are issued before any of them is consumed; slots past the
        // end of the row are filled with NaN, which every accumulator ignores (nansum semantics)
        for (int i0 = 0; i0 < nvec; i0 += 32 * RED_UNROLL) {
            float4 q[RED_UNROLL];
#pragma unroll
            for (int j = 0; j < RED_UNROLL; ++j) {
                const int i = i0 + j * 32 + lane;
                q[j] = i < nvec ? __ldcs(r4 + i) : make_float4(NAN, NAN, NAN, NAN);   // streaming: each byte is read once
            }
#pragma unroll
            for (int j = 0; j < RED_UNROLL; ++j) {
                if (MODE == 0) {
                    a.sum.x += (q[j].x == q[j].x ? q[j].x : 0.f) + (q[j].y == q[j].y ? q[j].y : 0.f)
                             + (q[j].z == q[j].z ? q[j].z : 0.f) + (q[j].w == q[j].w ? q[j].w : 0.f);
                } else {
                    const int pad = (i0 + j * 32 + lane) < nvec ? 0 : 4;          // NaN fillers are not data
                    snr_take4(a, q[j]);
                    a.nan -= pad;
                }
            }
        }
        for (int z = head + 4 * nvec + lane; z < nz; z += 32) {
            float v = row[z];
            if (MODE == 0) { if (v == v) a.sum.x += v; } else snr_take1(a.sum.x, a.asum.x, a.mx, a.nan, v);
        }
        float sum = (a.sum.x + a.sum.y) + (a.sum.z + a.sum.w), asum = (a.asum.x + a.asum.y) + (a.asum.z + a.asum.w);
        int nan = a.nan;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (MODE == 1) {
                asum += __shfl_xor_sync(0xffffffffu, asum, o);
                a.mx = fmaxf(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, o));
                n1 += __shfl_xor_sync(0xffffffffu, n1, o);
                n2 += __shfl_xor_sync(0xffffffffu, n2, o);
                nan += __shfl_xor_sync(0xffffffffu, nan, o);
                ncnt += __shfl_xor_sync(0xffffffffu, ncnt, o);
            }
        }
        if (lane == 0) {
            if (deep) deep[s] = s < n_live ? sum : 0.0f;
            if (MODE == 1) {
                const int cnt = nz - nan;
                float mean = sum / (float)cnt;
                double mu = ncnt > 0 ? n1 / ncnt : 0.0;
                float var = ncnt > 0 ? (float)fmax(n2 / ncnt - mu * mu, 0.0) : NAN;
                float std_out = sqrtf(var);
                float v = (a.mx - mean) / sqrtf(fabsf(std_out));          // LuciBase.py:1159-1161
                if (v < 0.0f) v = 0.0f;                                   // :1162-1163
                else v = v / sqrtf(asum / (float)cnt);                    // :1165
                snr[s] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------ TMA-staged streaming form
// The same reductions with the spectra staged through shared memory by the bulk-copy engine (cp.async.bulk, SASS
// UBLKCP) instead of per-lane loads: a CTA walks tiles of RED_TILE_ROWS consecutive spaxels (contiguous in the cube,
// a multiple of 16 bytes), keeps RED_STAGES tiles in flight behind mbarriers, and its 8 warps each reduce one row of
// the tile that has landed.  The bytes in flight per SM (4 CTAs x 2 tiles x 27 KB at nz = 841) no longer depend on
// the register budget of the arithmetic, which is what the SNR map needs to stay HBM-bound; two stages rather than
// three because four resident CTAs (32 warps) hide the per-tile tail (shuffles, scalar epilogue, barrier) better
// than a deeper ring does (measured: SNR 0.70 -> 0.91 of the HBM copy peak).
#define RED_TILE_ROWS 8
#ifndef RED_STAGES
#define RED_STAGES 2
#endif
#define RED_TMA_MAX_TILE_BYTES (40 * 1024)

__device__ __forceinline__ unsigned red_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void red_mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(red_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void red_mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(red_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(red_smem_u32(dst)), "l"(src), "r"(bytes), "r"(red_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool red_mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(red_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

template <int MODE>
__global__ void __launch_bounds__(RED_WARPS * 32)
lf_reduce_tma_kernel(const float* __restrict__ cube, long long n_tiles, int nz, long long n_live, int nlo, int nhi,
                     float* __restrict__ snr, float* __restrict__ deep)
{
    extern __shared__ __align__(128) unsigned char red_smem[];
    __shared__ __align__(8) unsigned long long full_bar[RED_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tile_bytes = (unsigned)RED_TILE_ROWS * (unsigned)nz * 4u;
    const unsigned stage_stride = (tile_bytes + 127u) & ~127u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < RED_STAGES; ++i) red_mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long first = blockIdx.x, step = gridDim.x;
    if (threadIdx.x == 0) {                                          // prologue: fill the ring
        for (int i = 0; i < RED_STAGES; ++i) {
            const long long t = first + (long long)i * step;
            if (t < n_tiles) {
                red_mbar_expect_tx(&full_bar[i], tile_bytes);
                red_bulk_g2s(red_smem + (size_t)i * stage_stride, cube + t * (long long)RED_TILE_ROWS * nz, tile_bytes, &full_bar[i]);
            }
        }
    }
    int stage = 0;
    unsigned parity = 0;
    for (long long t = first; t < n_tiles; t += step) {
        long long spins = 0;
        while (!red_mbar_try_wait(&full_bar[stage], parity)) { if (++spins > (1ll << 26)) __trap(); }
        const float* row = reinterpret_cast<const float*>(red_smem + (size_t)stage * stage_stride) + warp * nz;
        const long long s = t * RED_TILE_ROWS + warp;
        // Fast path: no NaN tests at all -- fmaxf drops NaN operands by itself and a NaN anywhere poisons `sum`, which
        // is checked once per row; only rows that do hold NaNs are redone with the per-sample nansum tests.
        float sum = 0.0f, asum = 0.0f, mx = -INFINITY;
        int nan = 0;
        {
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            float m0 = -INFINITY, m1 = -INFINITY;
            // 16-byte shared loads over the aligned body of the row (the row starts 4-byte aligned inside the tile)
            int head = (int)(((16 - ((uintptr_t)row & 15)) & 15) >> 2);
            if (head > nz) head = nz;
            if (lane < head) { const float v = row[lane]; s0 += v; if (MODE == 1) { a0 += fabsf(v); m0 = fmaxf(m0, v); } }
            const int nvec = (nz - head) >> 2;
            const float4* r4 = reinterpret_cast<const float4*>(row + head);
            for (int i = lane; i < nvec; i += 32) {
                const float4 q = r4[i];
                s0 += q.x; s1 += q.y; s2 += q.z; s3 += q.w;
                if (MODE == 1) {
                    a0 += fabsf(q.x); a1 += fabsf(q.y); a2 += fabsf(q.z); a3 += fabsf(q.w);
                    m0 = fmaxf(m0, fmaxf(q.x, q.y)); m1 = fmaxf(m1, fmaxf(q.z, q.w));
                }
            }
            for (int z = head + 4 * nvec + lane; z < nz; z += 32) {
                const float v = row[z]; s0 += v; if (MODE == 1) { a0 += fabsf(v); m0 = fmaxf(m0, v); }
            }
            sum = (s0 + s1) + (s2 + s3); asum = (a0 + a1) + (a2 + a3); mx = fmaxf(m0, m1);
        }
        if (__any_sync(0xffffffffu, sum != sum)) {                    // the row holds NaNs: nansum semantics, per sample
            sum = 0.0f; asum = 0.0f; mx = -INFINITY; nan = 0;
            for (int z = lane; z < nz; z += 32) {
                const float v = row[z];
                if (MODE == 0) sum += (v == v ? v : 0.0f); else snr_take1(sum, asum, mx, nan, v);
            }
        }
        // noise-window variance: two moments about the window's first sample (a shared-memory broadcast here), which
        // keeps FP32 stable for any continuum level
        float n1 = 0.0f, n2 = 0.0f;
        int ncnt = 0;
        if (MODE == 1) {
            float shift = row[nlo < nz ? nlo : 0];
            if (shift != shift) shift = 0.0f;
            for (int zz = nlo + lane; zz < nhi; zz += 32) {
                const float v = row[zz];
                if (v == v) { const float d = v - shift; n1 += d; n2 += d * d; ncnt += 1; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (MODE == 1) {
                asum += __shfl_xor_sync(0xffffffffu, asum, o);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                n1 += __shfl_xor_sync(0xffffffffu, n1, o);
                n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            }
        }
        if (MODE == 1) { nan = __reduce_add_sync(0xffffffffu, nan); ncnt = __reduce_add_sync(0xffffffffu, ncnt); }
        if (lane == 0) {
            if (deep) deep[s] = s < n_live ? sum : 0.0f;
            if (MODE == 1) {
                const int cnt = nz - nan;
                float mean = sum / (float)cnt;
                float mu = ncnt > 0 ? n1 / ncnt : 0.0f;
                float var = ncnt > 0 ? fmaxf(n2 / ncnt - mu * mu, 0.0f) : NAN;
                float std_out = sqrtf(var);
                float v = (mx - mean) / sqrtf(fabsf(std_out));            // LuciBase.py:1159-1161
                if (v < 0.0f) v = 0.0f;                                   // :1162-1163
                else v = v / sqrtf(asum / (float)cnt);                    // :1165
                snr[s] = v;
            }
        }
        __syncthreads();                                             // every warp is done with this stage
        if (threadIdx.x == 0) {
            const long long tn = t + (long long)RED_STAGES * step;
            if (tn < n_tiles) {
                red_mbar_expect_tx(&full_bar[stage], tile_bytes);
                red_bulk_g2s(red_smem + (size_t)stage * stage_stride, cube + tn * (long long)RED_TILE_ROWS * nz, tile_bytes, &full_bar[stage]);
            }
        }
        if (++stage == RED_STAGES) { stage = 0; parity ^= 1u; }
    }
}

// SNR method 2: flux-window sum minus n * min(sigma_clip(sky, sigma=1, maxiters=10)), over std of the noise
// window (LuciBase.py:1144-1157, 1167-1169).  sigma_clip: centre = median, spread = population std.
// The finite samples are sorted ONCE in shared memory (bitonic, padded with +inf): a clip keeps the values inside an
// interval, so the survivors are always a contiguous range [a, a+m) of the sorted spectrum -- the median is an index,
// the new range two binary searches, the clipped minimum is v[a].
__device__ __forceinline__ void warp_bitonic_sort(float* v, int npad, int lane)
{
    for (int k = 2; k <= npad; k <<= 1) {
        for (int lj = 31 - __clz(k >> 1); lj >= 0; --lj) {               // partner distance j = 1 << lj
            const int j = 1 << lj;
            for (int t = lane; t < (npad >> 1); t += 32) {
                const int i = ((t >> lj) << (lj + 1)) | (t & (j - 1)), l = i + j;   // i has bit lj clear
                const bool up = (i & k) == 0;
                const float a = v[i], b = v[l];
                const float lo = fminf(a, b), hi = fmaxf(a, b);
                v[i] = up ? lo : hi; v[l] = up ? hi : lo;
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(RED_WARPS * 32)
lf_snr2_kernel(const float* __restrict__ cube, long long n_spaxel, int nz, int npad, long long n_live, int flo, int fhi,
               int nlo, int nhi, float* __restrict__ snr, float* __restrict__ deep)
{
    extern __shared__ __align__(16) unsigned char smem2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* v = reinterpret_cast<float*>(smem2) + (size_t)warp * npad;
    const long long stride = (long long)gridDim.x * RED_WARPS;
    for (long long s = (long long)blockIdx.x * RED_WARPS + warp; s < n_spaxel; s += stride) {
        const float* row = cube + s * (long long)nz;
        float sum = 0.f, fsum = 0.f; double n1 = 0.0, n2 = 0.0; int ncnt = 0, m = 0;
        float allmin = INFINITY;
        __syncwarp();
        for (int z = lane; z < npad; z += 32) {
            float x = z < nz ? __ldcs(row + z) : NAN;
            const bool fin = (x == x) && fabsf(x) != INFINITY;
            v[z] = fin ? x : INFINITY;                                    // non-finite samples sort to the end
            m += fin;
            if (x == x) { sum += x; allmin = fminf(allmin, x); if (z >= flo && z < fhi) fsum += x;
                          if (z >= nlo && z < nhi) { n1 += x; n2 += (double)x * x; ++ncnt; } }
        }
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o); fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
            n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            ncnt += __shfl_xor_sync(0xffffffffu, ncnt, o); m += __shfl_xor_sync(0xffffffffu, m, o);
            allmin = fminf(allmin, __shfl_xor_sync(0xffffffffu, allmin, o));
        }
        warp_bitonic_sort(v, npad, lane);
        // sigma_clip(sky, sigma=1, maxiters=10): median / population std of the survivors v[a .. a+m)
        int a = 0;
        for (int it = 0; it < 10 && m > 0; ++it) {
            const double med = (m & 1) ? (double)v[a + m / 2] : 0.5 * ((double)v[a + m / 2 - 1] + (double)v[a + m / 2]);
            double s1 = 0.0, s2 = 0.0;
            for (int z = a + lane; z < a + m; z += 32) { double d = (double)v[z] - med; s1 += d; s2 += d * d; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            const double mu = s1 / m;
            const double sd = sqrt(fmax(s2 / m - mu * mu, 0.0));
            const double lo_b = med - sd, hi_b = med + sd;
            // survivors: lo_b <= v <= hi_b, a contiguous sub-range (binary searches, every lane the same)
            int l0 = a, l1 = a + m;                                       // first index with v >= lo_b
            while (l0 < l1) { int mid = (l0 + l1) >> 1; if ((double)v[mid] >= lo_b) l1 = mid; else l0 = mid + 1; }
            int h0 = l0, h1 = a + m;                                      // first index with v > hi_b
            while (h0 < h1) { int mid = (h0 + h1) >> 1; if ((double)v[mid] > hi_b) h1 = mid; else h0 = mid + 1; }
            const int removed = m - (h0 - l0);
            a = l0; m = h0 - l0;
            if (removed == 0) break;
        }
        const float cmin = m > 0 ? v[a] : allmin;                         // LuciBase.py:1151-1152
        if (lane == 0) {
            if (deep) deep[s] = s < n_live ? sum : 0.0f;
            double mu = ncnt > 0 ? n1 / ncnt : 0.0;
            double sd = ncnt > 0 ? sqrt(fmax(n2 / ncnt - mu * mu, 0.0)) : NAN;
            double val = ((double)fsum - (double)cmin * (double)(fhi - flo)) / sd;   // :1153, 1167
            snr[s] = val < 0.0 ? 0.0f : (float)val;
        }
    }
}

// ---- SNR method 2, register form (spectra of up to 32 NR samples) ----------------------------------------------
// The clip of LuciBase.py:1146-1152 needs, ten times over, the median / mean / std of the samples inside a shrinking value
// interval and finally their minimum.  The spectrum is sorted ONCE, in registers: element e of the padded spectrum lives in
// lane e / NR, register e % NR, so the 40 compare-exchange stages of the bitonic network whose partner distance is below NR
// are plain FMNMX pairs on the lane's own registers and only the 15 stages that cross lanes use shuffles (2.2 k warp
// instructions for 1024 keys; the shared-memory sort of lf_snr2_kernel spends ~9 k plus a barrier per stage).  The sorted
// keys then go to shared memory once (row stride NR + 1: conflict-free), and every clip iteration works on ranks: the
// survivors are a contiguous rank range [a, a + m), their float64 moments come from per-lane prefix sums plus the two
// partial lanes at the ends of the range, the new range from two ballot searches (lane heads, then one lane's row).
template <int NR>
__device__ __forceinline__ void lf_regsort(float (&v)[NR], int lane)
{
    // flip-merge form of the bitonic sorter (all comparators ascending): block size k: compare e with its mirror
    // e ^ (k - 1), then half-cleaners at distances k/4 ... 1
#pragma unroll
    for (int k = 2; k <= NR; k <<= 1) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int p = r ^ (k - 1);
            if (r < p) { const float lo = fminf(v[r], v[p]), hi = fmaxf(v[r], v[p]); v[r] = lo; v[p] = hi; }
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1)
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if ((r & j) == 0) { const float lo = fminf(v[r], v[r | j]), hi = fmaxf(v[r], v[r | j]); v[r] = lo; v[r | j] = hi; }
    }
#pragma unroll 1
    for (int m = 2; m <= 32; m <<= 1) {                      // blocks of m lanes (k = m NR keys)
        {
            const bool lower = (lane & (m >> 1)) == 0;
            float t[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) t[r] = __shfl_xor_sync(0xffffffffu, v[NR - 1 - r], m - 1);
#pragma unroll
            for (int r = 0; r < NR; ++r) v[r] = lower ? fminf(v[r], t[r]) : fmaxf(v[r], t[r]);
        }
#pragma unroll 1
        for (int jl = m >> 2; jl > 0; jl >>= 1) {
            const bool lower = (lane & jl) == 0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const float p = __shfl_xor_sync(0xffffffffu, v[r], jl);
                v[r] = lower ? fminf(v[r], p) : fmaxf(v[r], p);
            }
        }
#pragma unroll
        for (int j = NR >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if ((r & j) == 0) { const float lo = fminf(v[r], v[r | j]), hi = fmaxf(v[r], v[r | j]); v[r] = lo; v[r | j] = hi; }
    }
}

__device__ __forceinline__ double lf_warp_sum_d(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

template <int NR> struct LfSnr2Smem { enum { PER_WARP = 2 * 34 * 8 + 4 * 32 * (NR + 1) + 4 * 32 * NR }; };

template <int NR>
__global__ void __launch_bounds__(RED_WARPS * 32)
lf_snr2_reg_kernel(const float* __restrict__ cube, long long n_spaxel, int nz, long long n_live, int flo, int fhi,
                   int nlo, int nhi, float* __restrict__ snr, float* __restrict__ deep)
{
    constexpr int ROW = NR + 1;
    extern __shared__ __align__(16) unsigned char smem_snr2[];      // per warp: pre[2][34] doubles | sv[32 ROW] | nxt[32 NR]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem_snr2 + (size_t)warp * LfSnr2Smem<NR>::PER_WARP;
    double* pre1 = reinterpret_cast<double*>(mine);
    double* pre2 = pre1 + 34;
    float* sv = reinterpret_cast<float*>(pre2 + 34);
    float* nxt = sv + 32 * ROW;
    const long long stride = (long long)gridDim.x * RED_WARPS;
    // The sort and the clip of one spectrum keep a warp busy for ~10 k cycles without touching memory: the NEXT spectrum of
    // the warp is fetched meanwhile with cp.async into a per-warp staging row, so its DRAM latency is never waited for.
    auto prefetch = [&](long long sp) {
        if (sp < n_spaxel) {
            const float* row = cube + sp * (long long)nz;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int z = lane + 32 * i;
                if (z < nz) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(nxt + z);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(row + z) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch((long long)blockIdx.x * RED_WARPS + warp);
    for (long long s = (long long)blockIdx.x * RED_WARPS + warp; s < n_spaxel; s += stride) {
        float v[NR];
        float sum = 0.f, fsum = 0.f; double n1 = 0.0, n2 = 0.0; int ncnt = 0, m = 0;
        float allmin = INFINITY;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NR; ++i) {                       // sample z = lane + 32 i (any order sorts)
            const int z = lane + 32 * i;
            v[i] = z < nz ? nxt[z] : NAN;
        }
        __syncwarp();
        prefetch(s + stride);
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int z = lane + 32 * i;
            const float x = v[i];
            const bool nn = x == x;
            const bool fin = nn && fabsf(x) != INFINITY;
            m += fin;
            sum += nn ? x : 0.0f;
            allmin = fminf(allmin, x);                       // (fminf drops a NaN)
            fsum += (nn && z >= flo && z < fhi) ? x : 0.0f;
            if (nn && z >= nlo && z < nhi) { n1 += x; n2 += (double)x * x; ++ncnt; }
            v[i] = fin ? x : INFINITY;                       // non-finite samples sort to the end
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o); fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
            allmin = fminf(allmin, __shfl_xor_sync(0xffffffffu, allmin, o));
        }
        n1 = lf_warp_sum_d(n1); n2 = lf_warp_sum_d(n2);
        ncnt = __reduce_add_sync(0xffffffffu, ncnt); m = __reduce_add_sync(0xffffffffu, m);
        lf_regsort<NR>(v, lane);
        // sorted keys -> shared memory (rank e at sv[e + e / NR]); per-lane float64 moments about the first median
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NR; ++r) sv[lane * ROW + r] = v[r];
        __syncwarp();
        auto key = [&](int e) { return sv[e + e / NR]; };
        const double c0 = m > 0 ? (double)key(m / 2) : 0.0;
        {
            double l1 = 0.0, l2 = 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double d = (lane * NR + r < m) ? (double)v[r] - c0 : 0.0;
                l1 += d; l2 += d * d;
            }
            double i1 = l1, i2 = l2;                         // inclusive scan over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t1 = __shfl_up_sync(0xffffffffu, i1, o), t2 = __shfl_up_sync(0xffffffffu, i2, o);
                if (lane >= o) { i1 += t1; i2 += t2; }
            }
            pre1[lane + 1] = i1; pre2[lane + 1] = i2;       // pre[l] = moments of the lanes below l
            if (lane == 0) { pre1[0] = 0.0; pre2[0] = 0.0; }
        }
        __syncwarp();
        // moments of the ranks [0, e): whole lanes from the prefix, the partial lane read cooperatively
        auto partial = [&](int e, double& p1, double& p2) {
            const int l = e / NR, q = e - l * NR;
            double d = 0.0;
            if (lane < q) d = (double)sv[l * ROW + lane] - c0;     // (NR <= 32: one key per lane)
            p1 = d; p2 = d * d;
        };
        int a = 0;
        const float head = v[0];
        for (int it = 0; it < 10 && m > 0; ++it) {
            const double med = (m & 1) ? (double)key(a + m / 2) : 0.5 * ((double)key(a + m / 2 - 1) + (double)key(a + m / 2));
            double a1, a2, b1, b2;
            partial(a, a1, a2);
            partial(a + m, b1, b2);
            double s1 = lf_warp_sum_d(b1 - a1), s2 = lf_warp_sum_d(b2 - a2);
            s1 += pre1[(a + m) / NR] - pre1[a / NR];
            s2 += pre2[(a + m) / NR] - pre2[a / NR];
            // np.nanstd about the mean (moments about c0: a shift changes nothing).  A key survives when
            // |key - med| <= std, i.e. (key - med)^2 m^2 <= m s2 - s1^2: no square root, no division
            const double dm = (double)m;
            const double rhs = fmax(dm * s2 - s1 * s1, 0.0);
            auto is_below = [&](float k) { const double d = ((double)k - med) * dm; return d < 0.0 && d * d > rhs; };
            auto is_above = [&](float k) { const double d = ((double)k - med) * dm; return d > 0.0 && d * d > rhs; };
            // first rank that is not below / first rank that is above (the keys are sorted: two ballot searches each)
            int l0, h0;
            {
                const int nl = __popc(__ballot_sync(0xffffffffu, is_below(head)));     // lanes whose first key is below
                if (nl == 0) l0 = 0;
                else {
                    const unsigned in = __ballot_sync(0xffffffffu, lane < NR && is_below(sv[(nl - 1) * ROW + (lane < NR ? lane : 0)]));
                    l0 = (nl - 1) * NR + __popc(in);
                }
                const int nh = __popc(__ballot_sync(0xffffffffu, !is_above(head)));
                if (nh == 0) h0 = 0;
                else {
                    const unsigned in = __ballot_sync(0xffffffffu, lane < NR && !is_above(sv[(nh - 1) * ROW + (lane < NR ? lane : 0)]));
                    h0 = (nh - 1) * NR + __popc(in);
                }
            }
            if (l0 < a) l0 = a;
            if (h0 > a + m) h0 = a + m;
            if (h0 < l0) h0 = l0;
            const int removed = m - (h0 - l0);
            a = l0; m = h0 - l0;
            if (removed == 0) break;
        }
        const float cmin = m > 0 ? key(a) : allmin;                         // LuciBase.py:1151-1152
        if (lane == 0) {
            if (deep) deep[s] = s < n_live ? sum : 0.0f;
            double mu = ncnt > 0 ? n1 / ncnt : 0.0;
            double sd = ncnt > 0 ? sqrt(fmax(n2 / ncnt - mu * mu, 0.0)) : NAN;
            double val = ((double)fsum - (double)cmin * (double)(fhi - flo)) / sd;   // :1153, 1167
            snr[s] = val < 0.0 ? 0.0f : (float)val;
        }
        __syncwarp();
    }
}

// b x b nansum binning (bin_cube_function, LUCI/LuciUtility.py:332-341): one warp per OUTPUT spaxel walks the
// spectral axis (coalesced, no per-element index divisions), four output samples per lane in flight; the b*b input
// rows of a bin are read exactly once.
template <int B>                                        // B > 0: compile-time bin size (all B*B rows' loads in flight)
__global__ void __launch_bounds__(RED_WARPS * 32)
lf_bin_kernel(const float* __restrict__ cube, int ny, int nz, int b_rt, int x0, int y0, int nxb, int nyb,
              float* __restrict__ out)
{
    const int b = B > 0 ? B : b_rt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n_out = (long long)nxb * nyb;
    const long long stride = (long long)gridDim.x * RED_WARPS;
    for (long long sp = (long long)blockIdx.x * RED_WARPS + warp; sp < n_out; sp += stride) {
        const int j = (int)(sp % nyb), i = (int)(sp / nyb);
        const float* base = cube + ((long long)(x0 + i * b) * ny + (y0 + j * b)) * nz;
        float* orow = out + sp * (long long)nz;
        for (int z0 = lane; z0 < nz; z0 += 128) {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int di = 0; di < b; ++di)
#pragma unroll
                for (int dj = 0; dj < b; ++dj) {
                    const float* r = base + ((long long)di * ny + dj) * nz;
                    float q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int z = z0 + 32 * u; q[u] = z < nz ? __ldcs(r + z) : 0.0f; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] += (q[u] == q[u]) ? q[u] : 0.0f;
                }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int z = z0 + 32 * u; if (z < nz) orow[z] = acc[u]; }
        }
    }
}

// Synthetic cube: one warp per spaxel (LuciSim.Spectrum.create_spectrum, LuciSim.py:169-203).
struct LfSimArgs {
    int model, L, nz;
    double line_nm[LF_MAX_LINES];
    double axis_k;              // 1e7/(2 STEP (n_steps - zpd))
    const double* axis;         // [nz]
    long long n;
    const float* amp; const float* vel; const float* broad; const float* theta; const float* noise_sigma;
    const float* noise;
    float continuum, flux_scale;
    float* cube;
};

template <int MODEL>
__global__ void __launch_bounds__(RED_WARPS * 32) lf_sim_kernel(const __grid_constant__ LfSimArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LfWarp<32> w;
    const long long stride = (long long)gridDim.x * RED_WARPS;
    for (long long s = (long long)blockIdx.x * RED_WARPS + warp; s < a.n; s += stride) {
        double ct = cos((double)a.theta[s] * 0.017453292519943295);
        double corr = 1.0 / ct;
        float sinc_w = (float)(corr * a.axis_k);
        LfLine ln[LF_MAX_LINES];
        float amp[LF_MAX_LINES];
        const double x_ref = a.axis[0];
        for (int k = 0; k < a.L; ++k) {
            double lam_cm = 1e7 / a.line_nm[k];
            double pos = ((double)a.vel[s * a.L + k] / 3e5) * lam_cm + lam_cm;                 // LuciSim.py:185
            int idx = lf_nearest_axis(a.axis, a.nz, pos);                           // :186-187 (snap)
            double psn = a.axis[idx];
            double sigma = ((double)a.broad[s * a.L + k] * psn) / (3e5 * corr);                // :189
            float hi, lo;
            lf_split(psn - x_ref, hi, lo);
            lf_line_setup(MODEL, hi, lo, (float)sigma, sinc_w, ln[k]);
            amp[k] = a.amp[s * a.L + k];
        }
        const float piw = LF_PI_F / sinc_w, iw = 1.0f / sinc_w;
        float* row = a.cube + s * (long long)a.nz;
        const float* nrow = a.noise ? a.noise + s * (long long)a.nz : nullptr;
        float nsig = a.noise_sigma ? a.noise_sigma[s] : 0.0f;
        for (int z0 = 0; z0 < a.nz; z0 += 32) {
            int z = z0 + lane;
            bool live = z < a.nz;
            float x = live ? (float)(a.axis[z] - x_ref) : 0.0f;
            float m = a.continuum;
            for (int k = 0; k < a.L; ++k) {
                float dx = lf_dx(ln[k], x);
                bool far = false;
                if (MODEL == LF_SINCGAUSS) far = w.all(lf_is_far(ln[k], dx));
                float g, gp, gs;
                lf_profile<MODEL>(ln[k], dx, piw, iw, far, g, gp, gs);
                m += amp[k] * g;
            }
            if (live) {
                float nz_ = nrow ? nrow[z] : 0.0f;
                row[z] = a.flux_scale * (m + nsig * nz_);
            }
        }
    }
}

// Peak probes (roofline denominators that MEASURED_PEAKS.json does not carry): dependent-free FFMA chains and
// MUFU.EX2 chains, 8 independent accumulators per thread.
template <int MODE>
__global__ void __launch_bounds__(256) lf_peak_kernel(int iters, float* out)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <>
__global__ void __launch_bounds__(256) lf_peak_kernel<1>(int iters, float* out)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + .1f, a2 = a0 + .2f, a3 = a0 + .3f, a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
    for (int i = 0; i < iters; ++i) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
        a0 -= 1.0f; a1 -= 1.0f; a2 -= 1.0f; a3 -= 1.0f; a4 -= 1.0f; a5 -= 1.0f; a6 -= 1.0f; a7 -= 1.0f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" cudaError_t lf_run_peak(int mode, int iters, int blocks, float* out, cudaStream_t stream)
{
    if (mode == 0) lf_peak_kernel<0><<<blocks, 256, 0, stream>>>(iters, out);
    else lf_peak_kernel<1><<<blocks, 256, 0, stream>>>(iters, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------- host launchers
extern "C" cudaError_t lf_run_reduce(int mode, const float* cube, long long n_spaxel, int nz, long long n_live,
                                     int flo, int fhi, int nlo, int nhi, float* snr, float* deep, int sm_count,
                                     cudaStream_t stream)
{
    long long want = (n_spaxel + RED_WARPS - 1) / RED_WARPS;
    if (mode == 2 && nz <= 1024 && !getenv("LUCIFIT_SNR2_SMEM_SORT")) {
        // register sort: 16 keys per lane up to 512 samples, 32 up to 1024
        int occ = 1;
        const size_t smem = (size_t)RED_WARPS * (nz <= 512 ? (size_t)LfSnr2Smem<16>::PER_WARP : (size_t)LfSnr2Smem<32>::PER_WARP);
        cudaError_t e = nz <= 512 ? cudaFuncSetAttribute(lf_snr2_reg_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute(lf_snr2_reg_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (nz <= 512) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_snr2_reg_kernel<16>, RED_WARPS * 32, smem);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_snr2_reg_kernel<32>, RED_WARPS * 32, smem);
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        long long cap = (long long)sm_count * occ;
        int grid = (int)(want < cap ? want : cap); if (grid < 1) grid = 1;
        if (nz <= 512) lf_snr2_reg_kernel<16><<<grid, RED_WARPS * 32, smem, stream>>>(cube, n_spaxel, nz, n_live, flo, fhi, nlo, nhi, snr, deep);
        else lf_snr2_reg_kernel<32><<<grid, RED_WARPS * 32, smem, stream>>>(cube, n_spaxel, nz, n_live, flo, fhi, nlo, nhi, snr, deep);
        return cudaGetLastError();
    }
    if (mode == 2) {
        int npad = 32;
        while (npad < nz) npad <<= 1;
        size_t smem = (size_t)RED_WARPS * npad * sizeof(float);
        if (smem > 200 * 1024) return cudaErrorInvalidValue;      // spectra longer than 6400 samples
        cudaError_t e = cudaFuncSetAttribute(lf_snr2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_snr2_kernel, RED_WARPS * 32, smem);
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        long long cap = (long long)sm_count * occ;
        int grid = (int)(want < cap ? want : cap); if (grid < 1) grid = 1;
        lf_snr2_kernel<<<grid, RED_WARPS * 32, smem, stream>>>(cube, n_spaxel, nz, npad, n_live, flo, fhi, nlo, nhi, snr, deep);
        return cudaGetLastError();
    }
    // TMA-staged form for the whole tiles of RED_TILE_ROWS spaxels when the cube is 16-byte aligned and a stage ring
    // fits; the plain-load kernel below takes what is left (tail rows, unaligned views, very long spectra)
    const size_t tile_bytes = (size_t)RED_TILE_ROWS * nz * 4;
    if ((((uintptr_t)cube) & 15) == 0 && tile_bytes <= RED_TMA_MAX_TILE_BYTES && n_spaxel >= RED_TILE_ROWS && !getenv("LUCIFIT_NO_TMA")) {
        const long long n_tiles = n_spaxel / RED_TILE_ROWS;
        const size_t smem = RED_STAGES * ((tile_bytes + 127) & ~(size_t)127);
        cudaError_t e;
        int occ_t = 0;
        if (mode == 0) {
            e = cudaFuncSetAttribute(lf_reduce_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, lf_reduce_tma_kernel<0>, RED_WARPS * 32, smem);
        } else {
            e = cudaFuncSetAttribute(lf_reduce_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, lf_reduce_tma_kernel<1>, RED_WARPS * 32, smem);
        }
        if (e != cudaSuccess) return e;
        if (occ_t >= 1) {
            long long cap_t = (long long)sm_count * occ_t;
            int grid_t = (int)(n_tiles < cap_t ? n_tiles : cap_t);
            if (mode == 0) lf_reduce_tma_kernel<0><<<grid_t, RED_WARPS * 32, smem, stream>>>(cube, n_tiles, nz, n_live, 0, 0, nullptr, deep);
            else lf_reduce_tma_kernel<1><<<grid_t, RED_WARPS * 32, smem, stream>>>(cube, n_tiles, nz, n_live, nlo, nhi, snr, deep);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            const long long done = n_tiles * RED_TILE_ROWS;
            if (done == n_spaxel) return cudaSuccess;
            cube += done * (long long)nz; n_spaxel -= done; n_live -= done;
            if (snr) snr += done;
            if (deep) deep += done;
            want = (n_spaxel + RED_WARPS - 1) / RED_WARPS;
        }
    }
    int occ = 1;                                        // one resident wave, grid-stride inside
    if (mode == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_reduce_kernel<0>, RED_WARPS * 32, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_reduce_kernel<1>, RED_WARPS * 32, 0);
    if (occ < 1) occ = 1;
    long long cap = (long long)sm_count * occ;
    int grid = (int)(want < cap ? want : cap); if (grid < 1) grid = 1;
    if (mode == 0)
        lf_reduce_kernel<0><<<grid, RED_WARPS * 32, 0, stream>>>(cube, n_spaxel, nz, n_live, 0, 0, 0, 0, nullptr, deep);
    else
        lf_reduce_kernel<1><<<grid, RED_WARPS * 32, 0, stream>>>(cube, n_spaxel, nz, n_live, flo, fhi, nlo, nhi, snr, deep);
    return cudaGetLastError();
}

extern "C" cudaError_t lf_run_bin(const float* cube, int ny, int nz, int b, int x0, int y0, int nxb, int nyb,
                                  float* out, int sm_count, cudaStream_t stream)
{
    long long want = ((long long)nxb * nyb + RED_WARPS - 1) / RED_WARPS;
    int occ = 1;
    if (b == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_bin_kernel<2>, RED_WARPS * 32, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lf_bin_kernel<0>, RED_WARPS * 32, 0);
    if (occ < 1) occ = 1;
    long long cap = (long long)sm_count * occ;
    int grid = (int)(want < cap ? want : cap); if (grid < 1) grid = 1;
    if (b == 2) lf_bin_kernel<2><<<grid, RED_WARPS * 32, 0, stream>>>(cube, ny, nz, b, x0, y0, nxb, nyb, out);
    else lf_bin_kernel<0><<<grid, RED_WARPS * 32, 0, stream>>>(cube, ny, nz, b, x0, y0, nxb, nyb, out);
    return cudaGetLastError();
}

extern "C" cudaError_t lf_run_sim(const LfSimArgs* a, int sm_count, cudaStream_t stream)
{
    long long want = (a->n + RED_WARPS - 1) / RED_WARPS;
    long long cap = (long long)sm_count * 4;
    int grid = (int)(want < cap ? want : cap); if (grid < 1) grid = 1;
    if (a->model == 0) lf_sim_kernel<0><<<grid, RED_WARPS * 32, 0, stream>>>(*a);
    else if (a->model == 1) lf_sim_kernel<1><<<grid, RED_WARPS * 32, 0, stream>>>(*a);
    else lf_sim_kernel<2><<<grid, RED_WARPS * 32, 0, stream>>>(*a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ cube ingest
// Clamp rules of Luci.read_in_cube (LuciBase.py:123-129): non-finite samples, fluxes below `lo` (-1e-16) and above
// `hi` (1e-9) are replaced by `fill` (1e-22), in place.  One 16-byte load and store per thread and step; the scalar
// head / tail covers buffers that are not 16-byte aligned (views into a larger cube).
__global__ void __launch_bounds__(256) lf_sanitize_kernel(float* __restrict__ cube, long long n, float lo, float hi, float fill)
{
    const long long head = (long long)(((16 - ((uintptr_t)cube & 15)) & 15) / 4);
    const long long h = head < n ? head : n;
    const long long n4 = (n - h) / 4;
    float4* v = reinterpret_cast<float4*>(cube + h);
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    auto fix = [=](float x) { return (isfinite(x) && x >= lo && x <= hi) ? x : fill; };
    for (long long i = t0; i < n4; i += stride) {
        float4 q = v[i];
        q.x = fix(q.x); q.y = fix(q.y); q.z = fix(q.z); q.w = fix(q.w);
        v[i] = q;
    }
    if (t0 < h) cube[t0] = fix(cube[t0]);
    const long long tail = h + 4 * n4 + t0;
    if (tail < n) cube[tail] = fix(cube[tail]);
}

extern "C" cudaError_t lf_run_sanitize(float* cube, long long n, float lo, float hi, float fill, int sm_count, cudaStream_t stream)
{
    long long want = (n / 4 + 255) / 256;
    long long cap = (long long)sm_count * 8;
    int grid = (int)(want < cap ? want : cap); if (grid < 1) grid = 1;
    lf_sanitize_kernel<<<grid, 256, 0, stream>>>(cube, n, lo, hi, fill);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------- segmented sums
// Integrated spectra: sums[seg][z] += cube[row][z] over a list of (row, segment) pairs -- the spaxel loops of
// Luci.fit_spectrum_region (LuciBase.py:1015-1023), extract_spectrum(_region) (:872-889, :931-941), the per-bin
// fit_spectrum_region calls of fit_wvt (:1434-1447) and the region spectra of skyline_calibration (:1254-1256), in
// float64 like the reference.  A CTA owns a contiguous chunk of the list and the threads own spectral columns
// (coalesced row reads, LF_SEG_ROWS rows in flight); the running sums stay in registers while consecutive pairs name the
// same segment and are flushed with one float64 atomic per column when the segment changes or the chunk ends, so a
// list ordered by segment costs (segments + CTAs) x nz atomics in all.
#define LF_SEG_THREADS 256
#define LF_SEG_COLS 4
#define LF_SEG_ROWS 4
__global__ void __launch_bounds__(LF_SEG_THREADS)
lf_segsum_kernel(const float* __restrict__ cube, int nz, const long long* __restrict__ row_index,
                 const int* __restrict__ segment, long long n_sel, long long rows_per_cta, int nansum,
                 double* __restrict__ sums)
{
    const long long lo = (long long)blockIdx.x * rows_per_cta;
    const long long hi = lo + rows_per_cta < n_sel ? lo + rows_per_cta : n_sel;
    for (int z0 = threadIdx.x; z0 < nz; z0 += LF_SEG_THREADS * LF_SEG_COLS) {
        double acc[LF_SEG_COLS];
#pragma unroll
        for (int c = 0; c < LF_SEG_COLS; ++c) acc[c] = 0.0;
        int cur = -1;
        auto flush = [&]() {
            if (cur >= 0) {
#pragma unroll
                for (int c = 0; c < LF_SEG_COLS; ++c) {
                    const int z = z0 + c * LF_SEG_THREADS;
                    if (z < nz) atomicAdd(sums + (long long)cur * nz + z, acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < LF_SEG_COLS; ++c) acc[c] = 0.0;
        };
        for (long long i0 = lo; i0 < hi; i0 += LF_SEG_ROWS) {
            int seg[LF_SEG_ROWS];
            float q[LF_SEG_ROWS][LF_SEG_COLS];
#pragma unroll
            for (int r = 0; r < LF_SEG_ROWS; ++r) {
                const long long i = i0 + r;
                seg[r] = i < hi ? (segment ? segment[i] : 0) : -1;
                const long long row = i < hi ? (row_index ? row_index[i] : i) : 0;
                const float* rp = cube + row * (long long)nz;
#pragma unroll
                for (int c = 0; c < LF_SEG_COLS; ++c) {
                    const int z = z0 + c * LF_SEG_THREADS;
                    q[r][c] = (seg[r] >= 0 && z < nz) ? __ldcs(rp + z) : 0.0f;
                }
            }
#pragma unroll
            for (int r = 0; r < LF_SEG_ROWS; ++r) {
                if (seg[r] < 0) continue;                              // (uniform over the CTA)
                if (seg[r] != cur) { flush(); cur = seg[r]; }
#pragma unroll
                for (int c = 0; c < LF_SEG_COLS; ++c) {
                    const float v = q[r][c];
                    acc[c] += (nansum && v != v) ? 0.0 : (double)v;
                }
            }
        }
        flush();
    }
}

extern "C" cudaError_t lf_run_segsum(const float* cube, int nz, const long long* row_index, const int* segment,
                                     long long n_sel, int n_segments, int nansum, double* sums, int sm_count,
                                     cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)n_segments * nz, stream);
    if (e != cudaSuccess || n_sel <= 0) return e;
    long long ctas = (long long)sm_count * 8;
    long long per = (n_sel + ctas - 1) / ctas;
    if (per < 4 * LF_SEG_ROWS) per = 4 * LF_SEG_ROWS;                 // short lists: fewer CTAs, fewer atomics
    per = (per + LF_SEG_ROWS - 1) / LF_SEG_ROWS * LF_SEG_ROWS;
    const int grid = (int)((n_sel + per - 1) / per);
    lf_segsum_kernel<<<grid, LF_SEG_THREADS, 0, stream>>>(cube, nz, row_index, segment, n_sel, per, nansum, sums);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ output maps
// rows[n][K] (what lucifit_fit_batch writes: one row per fitted spaxel, map order) -> field-major maps[K][n], the layout
// of the 2-D maps Luci.fit_cube returns and LuciUtility.save_fits writes (LuciBase.py:531-551), and, in the same pass, a
// second copy with the bytes of every float reversed: FITS is big-endian, so the files are then written straight from
// that copy without a byte swap on the host.  A CTA moves 32 consecutive rows (one contiguous 32 K-float block) through
// a padded shared-memory tile; both sides are coalesced.
#define LF_PACK_ROWS 32
__global__ void __launch_bounds__(256)
lf_pack_maps_kernel(const float* __restrict__ rows, long long n, int K, float* __restrict__ maps, unsigned* __restrict__ maps_be)
{
    extern __shared__ float pack_tile[];                   // [LF_PACK_ROWS][K + 1]
    const int KP = K + 1;
    for (long long i0 = (long long)blockIdx.x * LF_PACK_ROWS; i0 < n; i0 += (long long)gridDim.x * LF_PACK_ROWS) {
        const int nr = (int)((n - i0) < LF_PACK_ROWS ? (n - i0) : LF_PACK_ROWS);
        const float* src = rows + i0 * K;
        for (int e = threadIdx.x; e < nr * K; e += blockDim.x) pack_tile[(e / K) * KP + (e % K)] = __ldcs(src + e);
        __syncthreads();
        for (int e = threadIdx.x; e < K * LF_PACK_ROWS; e += blockDim.x) {
            const int k = e / LF_PACK_ROWS, r = e % LF_PACK_ROWS;
            if (r < nr) {
                const float v = pack_tile[r * KP + k];
                maps[(long long)k * n + i0 + r] = v;
                if (maps_be) maps_be[(long long)k * n + i0 + r] = __byte_perm(__float_as_uint(v), 0, 0x0123);
            }
        }
        __syncthreads();
    }
}

extern "C" cudaError_t lf_run_pack_maps(const float* rows, long long n, int K, float* maps, unsigned* maps_be, int sm_count,
                                        cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    long long want = (n + LF_PACK_ROWS - 1) / LF_PACK_ROWS;
    long long cap = (long long)sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    lf_pack_maps_kernel<<<grid, 256, sizeof(float) * LF_PACK_ROWS * (K + 1), stream>>>(rows, n, K, maps, maps_be);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ row medians
// np.nanmedian(sky[lo:hi]) of every row (the continuum level the absorption template is scaled by, LuciBase.py:354-356):
// the same register sort as SNR method 2, one warp per row; the two middle keys are averaged in float64.
template <int NR>
__global__ void __launch_bounds__(RED_WARPS * 32)
lf_rowmedian_kernel(const float* __restrict__ cube, long long n_rows, int nz, int lo, int hi, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = (long long)gridDim.x * RED_WARPS;
    for (long long s = (long long)blockIdx.x * RED_WARPS + warp; s < n_rows; s += stride) {
        const float* row = cube + s * (long long)nz;
        float v[NR];
        int m = 0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int z = lo + lane + 32 * i;
            const float x = z < hi ? __ldcs(row + z) : NAN;
            m += (x == x);
            v[i] = (x == x) ? x : INFINITY;                  // NaN samples and padding sort to the end
        }
        m = __reduce_add_sync(0xffffffffu, m);
        lf_regsort<NR>(v, lane);
        auto key = [&](int e) {                              // rank e lives in lane e / NR, register e % NR
            float mine = 0.0f;
#pragma unroll
            for (int r = 0; r < NR; ++r) if (r == (e % NR)) mine = v[r];
            return __shfl_sync(0xffffffffu, mine, e / NR);
        };
        float med = NAN;
        if (m > 0) med = (float)(0.5 * ((double)key((m - 1) / 2) + (double)key(m / 2)));
        if (lane == 0) out[s] = med;
    }
}

extern "C" cudaError_t lf_run_rowmedian(const float* cube, long long n_rows, int nz, int lo, int hi, float* out, int sm_count,
                                        cudaStream_t stream)
{
    if (n_rows <= 0) return cudaSuccess;
    const int len = hi - lo;
    if (len > 1024) return cudaErrorInvalidValue;
    long long want = (n_rows + RED_WARPS - 1) / RED_WARPS;
    long long cap = (long long)sm_count * 4;
    const int grid = (int)(want < cap ? want : cap);
    if (len <= 512) lf_rowmedian_kernel<16><<<grid, RED_WARPS * 32, 0, stream>>>(cube, n_rows, nz, lo, hi, out);
    else lf_rowmedian_kernel<32><<<grid, RED_WARPS * 32, 0, stream>>>(cube, n_rows, nz, lo, hi, out);
    return cudaGetLastError();
}
