// lucifit_capi.cu -- the extern "C" entry points of include/lucifit.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "lucifit_host.hpp"
#include "lucifit_fit_kernel.cuh"
#include "lucifit_prior.cuh"

// Address ranges mapped by lucifit_peer_open (base, bytes): lucifit_fit_stages asks whether its out_rows lies in one.
static std::mutex lf_peer_mu;
static std::vector<std::pair<uintptr_t, size_t> > lf_peer_ranges;
static bool lf_peer_mapped(const void* p)
{
    std::lock_guard<std::mutex> g(lf_peer_mu);
    const uintptr_t a = (uintptr_t)p;
    for (size_t i = 0; i < lf_peer_ranges.size(); ++i)
        if (a >= lf_peer_ranges[i].first && a - lf_peer_ranges[i].first < lf_peer_ranges[i].second) return true;
    return false;
}

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what)
{
    return fail(LUCIFIT_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#if defined(LF_DEV_MODEL) && defined(LF_DEV_L)
// developer build (csrc/dev_variant.sh): a single (model, n_lines) instance, for quick kernel experiments
#define LF_DEV_CAT2(a, b, c, d) a##b##c##d
#define LF_DEV_CAT(a, b, c, d) LF_DEV_CAT2(a, b, c, d)
extern "C" lf_launch_fn LF_DEV_CAT(lf_launcher_m, LF_DEV_MODEL, _l, LF_DEV_L)(int bucket);
static lf_launch_fn find_launcher(int model, int L, int bucket)
{
    if (model == LF_DEV_MODEL && L == LF_DEV_L) return LF_DEV_CAT(lf_launcher_m, LF_DEV_MODEL, _l, LF_DEV_L)(bucket);
    return nullptr;
}
#else
#define LF_DECL(m, l) extern "C" lf_launch_fn lf_launcher_m##m##_l##l(int bucket);
#define LF_DECL_ALL(m) LF_DECL(m, 1) LF_DECL(m, 2) LF_DECL(m, 3) LF_DECL(m, 4) LF_DECL(m, 5) LF_DECL(m, 6) LF_DECL(m, 7) LF_DECL(m, 8)
LF_DECL_ALL(0) LF_DECL_ALL(1) LF_DECL_ALL(2)

static lf_launch_fn find_launcher(int model, int L, int bucket)
{
#define LF_CASE(m, l) if (model == m && L == l) return lf_launcher_m##m##_l##l(bucket);
#define LF_CASE_ALL(m) LF_CASE(m, 1) LF_CASE(m, 2) LF_CASE(m, 3) LF_CASE(m, 4) LF_CASE(m, 5) LF_CASE(m, 6) LF_CASE(m, 7) LF_CASE(m, 8)
    LF_CASE_ALL(0) LF_CASE_ALL(1) LF_CASE_ALL(2)
    return nullptr;
}
#endif

struct LfSimArgs {
    int model, L, nz;
    double line_nm[LF_MAX_LINES];
    double axis_k;
    const double* axis;
    long long n;
    const float* amp; const float* vel; const float* broad; const float* theta; const float* noise_sigma;
    const float* noise;
    float continuum, flux_scale;
    float* cube;
};
extern "C" cudaError_t lf_run_reduce(int mode, const float* cube, long long n_spaxel, int nz, long long n_live,
                                     int flo, int fhi, int nlo, int nhi, float* snr, float* deep, int sm_count,
                                     cudaStream_t stream);
extern "C" cudaError_t lf_run_bin(const float* cube, int ny, int nz, int b, int x0, int y0, int nxb, int nyb,
                                  float* out, int sm_count, cudaStream_t stream);
extern "C" cudaError_t lf_run_sim(const LfSimArgs* a, int sm_count, cudaStream_t stream);
extern "C" cudaError_t lf_run_peak(int mode, int iters, int blocks, float* out, cudaStream_t stream);
extern "C" cudaError_t lf_run_rowmedian(const float* cube, long long n_rows, int nz, int lo, int hi, float* out, int sm_count,
                                        cudaStream_t stream);
extern "C" cudaError_t lf_run_pack_maps(const float* rows, long long n, int K, float* maps, unsigned* maps_be, int sm_count,
                                        cudaStream_t stream);
extern "C" cudaError_t lf_run_segsum(const float* cube, int nz, const long long* row_index, const int* segment,
                                     long long n_sel, int n_segments, int nansum, double* sums, int sm_count,
                                     cudaStream_t stream);
extern "C" cudaError_t lf_run_sanitize(float* cube, long long n, float lo, float hi, float fill, int sm_count, cudaStream_t stream);
extern "C" cudaError_t lf_run_prior(const LfPriorCfg* c, long long n, const float* cube, const long long* index,
                                    const float* theta, float* vel, float* broad, int sm_count, cudaStream_t stream);

static int sm_count_of_current_device(int* out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    static thread_local int cached_dev = -1, cached_sm = 0;
    if (cached_dev != dev) {
        e = cudaDeviceGetAttribute(&cached_sm, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
        cached_dev = dev;
    }
    *out = cached_sm;
    return 0;
}

extern "C" {

const char* lucifit_last_error(void) { return g_err.c_str(); }
const char* lucifit_version(void) { return "lucifit 0.1.0 (sm_100a)"; }

int lucifit_row_floats(int n_lines) { return 7 * n_lines + 5; }

static size_t lf_slice_of(int64_t n_spaxel)
{
    return (size_t)(n_spaxel < LF_SLICE_SPAXELS ? (n_spaxel > 0 ? n_spaxel : 1) : LF_SLICE_SPAXELS);
}

int lucifit_workspace_bytes(const lucifit_config* cfg, int64_t n_spaxel, size_t* bytes)
{
    if (!bytes) return fail(LUCIFIT_EINVAL, "bytes is NULL");
    LfPlan pl;
    std::string err = lf_make_plan(cfg, pl);
    if (!err.empty()) return fail(LUCIFIT_EINVAL, err);
    // constant tables + the state records of one slice (the batch is processed LF_SLICE_SPAXELS at a time)
    *bytes = pl.table_bytes + lf_align(lf_slice_of(n_spaxel) * pl.state_bytes + 16, 256);
    return 0;
}

int lucifit_fit_batch(const lucifit_config* cfg, int64_t n_spaxel, const float* cube,
                      const int64_t* spaxel_index, const float* theta_deg, const float* vel0, const float* broad0,
                      const float* bkg, float* out_rows, float* out_fit_sol, float* out_unc, int32_t* out_status,
                      void* workspace, size_t workspace_bytes, void* stream)
{
    return lucifit_fit_stages(cfg, LUCIFIT_STAGE_ALL, n_spaxel, cube, spaxel_index, theta_deg, vel0, broad0, bkg, out_rows,
                              out_fit_sol, out_unc, out_status, workspace, workspace_bytes, stream);
}

int lucifit_fit_stages(const lucifit_config* cfg, int stages, int64_t n_spaxel, const float* cube,
                       const int64_t* spaxel_index, const float* theta_deg, const float* vel0, const float* broad0,
                       const float* bkg, float* out_rows, float* out_fit_sol, float* out_unc, int32_t* out_status,
                       void* workspace, size_t workspace_bytes, void* stream)
{
    if ((stages & ~LUCIFIT_STAGE_ALL) != 0 || stages == 0) return fail(LUCIFIT_EINVAL, "stages must be a non-empty subset of LUCIFIT_STAGE_ALL");
    if (stages != LUCIFIT_STAGE_ALL && n_spaxel > LF_SLICE_SPAXELS)
        return fail(LUCIFIT_EINVAL, "a partial pipeline keeps its state in the workspace: at most 2^20 spaxels per call");
    LfPlan pl;
    std::string err = lf_make_plan(cfg, pl);
    if (!err.empty()) return fail(LUCIFIT_EINVAL, err);
    if (n_spaxel < 0) return fail(LUCIFIT_EINVAL, "n_spaxel < 0");
    if (n_spaxel == 0) return 0;
    if (!cube || !theta_deg || !out_rows || !out_status) return fail(LUCIFIT_EINVAL, "cube, theta_deg, out_rows and out_status are required");
    const size_t slice = lf_slice_of(n_spaxel);
    const size_t need = pl.table_bytes + lf_align(slice * pl.state_bytes + 16, 256);
    if (!workspace || workspace_bytes < need) return fail(LUCIFIT_EWORKSPACE, "workspace too small: call lucifit_workspace_bytes");
    if (((uintptr_t)workspace & 255) != 0) return fail(LUCIFIT_EINVAL, "workspace must be 256-byte aligned");
    lf_launch_fn fn = find_launcher(cfg->model, pl.L, pl.nv_t);
    if (!fn) return fail(LUCIFIT_EINVAL, "this (n_lines, tie pattern) combination is not instantiated in this build");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    cudaError_t e;
    e = cudaMemcpyAsync(ws + pl.off_axis, pl.axis.data(), sizeof(double) * pl.axis.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "copy axis");
    e = cudaMemcpyAsync(ws + pl.off_xo, pl.xo.data(), sizeof(float) * pl.xo.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "copy xo");
    if (!pl.trans.empty()) {
        e = cudaMemcpyAsync(ws + pl.off_trans, pl.trans.data(), sizeof(float) * pl.trans.size(), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return cuda_fail(e, "copy transmission");
    }
    // the staging vectors live in `pl` (pageable host memory): cudaMemcpyAsync from pageable memory returns once
    // the source has been staged, so `pl` may be destroyed when this function returns.
    LfLaunch a;
    a.cfg = pl.dev;
    a.ly = pl.ly;
    a.cfg.axis = (const double*)(ws + pl.off_axis);
    a.cfg.xo = (const float*)(ws + pl.off_xo);
    a.cfg.trans = pl.trans.empty() ? nullptr : (const float*)(ws + pl.off_trans);
    a.cfg.bkg = bkg;
    a.cube = cube;
    a.xo_bytes = (int)lf_align(sizeof(float) * pl.nfit, 16);
    a.stream = st;
    a.state = (float*)(ws + pl.table_bytes);
    a.unc_ws = pl.dev.uncertainty ? (double*)(ws + pl.table_bytes + lf_align(slice * sizeof(float) * pl.dev.state_floats, 16)) : nullptr;
    int rc = sm_count_of_current_device(&a.sm_count);
    if (rc) return rc;
    // rows that land in another process's / GPU's memory (a range this library mapped in lucifit_peer_open) leave the output
    // kernel as whole coalesced stores (cudaPointerGetAttributes cannot tell: it reports an IPC mapping as local memory)
    a.stage_rows = lf_peer_mapped(out_rows) ? 1 : 0;
    if (const char* ev = getenv("LUCIFIT_STAGE_ROWS")) a.stage_rows = atoi(ev) != 0;          // measurement override
    if (getenv("LUCIFIT_DEBUG")) fprintf(stderr, "lucifit: fit_stages n=%lld stage_rows=%d\n", (long long)n_spaxel, a.stage_rows);
    const size_t smem_max = (size_t)a.xo_bytes + (size_t)8 * pl.ly.prep_bytes;          // the smallest CTA lf_pick_warps falls back to
    if (smem_max > LF_SMEM_BUDGET) return fail(LUCIFIT_EINVAL, "spectrum too long for the shared-memory layout (nz * 8 warps > 200 KB)");
    const int K = 7 * pl.L + 5, PF = 3 * pl.L + 1;
    for (int64_t s0 = 0; s0 < n_spaxel; s0 += (int64_t)slice) {
        const int64_t ns = n_spaxel - s0 < (int64_t)slice ? n_spaxel - s0 : (int64_t)slice;
        a.n = ns;
        a.index = spaxel_index ? (const long long*)spaxel_index + s0 : nullptr;
        a.index_base = s0;
        a.theta = theta_deg + s0;
        a.vel0 = vel0 ? vel0 + s0 * (pl.dev.prior_groups ? pl.nv : 1) : nullptr;
        a.broad0 = broad0 ? broad0 + s0 * (pl.dev.prior_groups ? pl.ns : 1) : nullptr;
        a.out_rows = out_rows + s0 * K;
        a.out_sol = out_fit_sol ? out_fit_sol + s0 * PF : nullptr;
        a.out_unc = out_unc ? out_unc + s0 * PF : nullptr;
        a.status = out_status + s0;
        e = fn(a, stages);
        if (e != cudaSuccess) return cuda_fail(e, "launch fit stage kernels");
    }
    return 0;
}

int lucifit_deep_image(const float* cube, int64_t n_spaxel, int nz, int64_t n_zero_tail, float* deep, void* stream)
{
    if (!cube || !deep || n_spaxel < 0 || nz < 1 || n_zero_tail < 0) return fail(LUCIFIT_EINVAL, "bad argument");
    if (n_spaxel == 0) return 0;
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_reduce(0, cube, n_spaxel, nz, n_spaxel - n_zero_tail, 0, 0, 0, 0, nullptr, deep, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_reduce_kernel");
    return 0;
}

int lucifit_snr_map(const float* cube, int64_t n_spaxel, int nz, int method, int flux_lo, int flux_hi,
                    int noise_lo, int noise_hi, float* snr, float* deep, void* stream)
{
    if (!cube || !snr || n_spaxel < 0 || nz < 1) return fail(LUCIFIT_EINVAL, "bad argument");
    if (method != 1 && method != 2) return fail(LUCIFIT_EINVAL, "method must be 1 or 2");
    if (flux_lo < 0 || flux_hi > nz || noise_lo < 0 || noise_hi > nz) return fail(LUCIFIT_EINVAL, "window outside [0, nz]");
    if (n_spaxel == 0) return 0;
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_reduce(method, cube, n_spaxel, nz, n_spaxel, flux_lo, flux_hi, noise_lo, noise_hi, snr, deep, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch snr kernel");
    return 0;
}

int lucifit_bin_cube(const float* cube, int nx, int ny, int nz, int b, int x0, int x1, int y0, int y1,
                     float* binned, void* stream)
{
    if (!cube || !binned || b < 1 || nx < 1 || ny < 1 || nz < 1) return fail(LUCIFIT_EINVAL, "bad argument");
    if (x0 < 0 || x1 > nx || y0 < 0 || y1 > ny || x1 < x0 || y1 < y0) return fail(LUCIFIT_EINVAL, "region outside the cube");
    int nxb = (x1 - x0) / b, nyb = (y1 - y0) / b;
    if (nxb == 0 || nyb == 0) return 0;
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_bin(cube, ny, nz, b, x0, y0, nxb, nyb, binned, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_bin_kernel");
    return 0;
}

int lucifit_sim_cube(const lucifit_config* cfg, int64_t n_spaxel, const float* amp, const float* vel,
                     const float* broad, const float* theta_deg, const float* noise_sigma, const float* noise,
                     float continuum, float flux_scale, float* cube, void* workspace, size_t workspace_bytes,
                     void* stream)
{
    if (!cfg || !cfg->axis || cfg->nz < 8) return fail(LUCIFIT_EINVAL, "config/axis missing");
    if (cfg->n_lines < 1 || cfg->n_lines > LF_MAX_LINES) return fail(LUCIFIT_EINVAL, "n_lines must be in [1, 8]");
    if (cfg->model < 0 || cfg->model > 2) return fail(LUCIFIT_EINVAL, "bad model");
    if (!amp || !vel || !broad || !theta_deg || !cube) return fail(LUCIFIT_EINVAL, "amp, vel, broad, theta_deg and cube are required");
    size_t need = lf_align(sizeof(double) * cfg->nz, 256);
    if (!workspace || workspace_bytes < need) return fail(LUCIFIT_EWORKSPACE, "workspace too small (need 8*nz bytes, 256-aligned)");
    if (n_spaxel <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(workspace, cfg->axis, sizeof(double) * cfg->nz, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "copy axis");
    LfSimArgs a;
    a.model = cfg->model; a.L = cfg->n_lines; a.nz = cfg->nz;
    for (int k = 0; k < LF_MAX_LINES; ++k) a.line_nm[k] = k < cfg->n_lines ? cfg->line_nm[k] : 1.0;
    a.axis_k = 1e7 / (2.0 * cfg->step_nm * (cfg->n_steps - cfg->zpd_index));
    a.axis = (const double*)workspace;
    a.n = n_spaxel; a.amp = amp; a.vel = vel; a.broad = broad; a.theta = theta_deg; a.noise_sigma = noise_sigma;
    a.noise = noise; a.continuum = continuum; a.flux_scale = flux_scale; a.cube = cube;
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    e = lf_run_sim(&a, sm, st);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_sim_kernel");
    return 0;
}

int lucifit_estimate_prior(const lucifit_config* cfg, int64_t n_spaxel, const float* cube, const int64_t* spaxel_index,
                           const float* theta_deg, const float* bkg, double v_min, double v_max, double v_step,
                           float* vel0, float* broad0, void* stream)
{
    LfPriorCfg pc;
    std::string err = lf_make_prior_cfg(cfg, v_min, v_max, v_step, pc);
    if (!err.empty()) return fail(LUCIFIT_EINVAL, err);
    if (n_spaxel < 0) return fail(LUCIFIT_EINVAL, "n_spaxel < 0");
    if (n_spaxel == 0) return 0;
    if (!cube || !theta_deg || !vel0 || !broad0) return fail(LUCIFIT_EINVAL, "cube, theta_deg, vel0 and broad0 are required");
    pc.bkg = bkg;
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_prior(&pc, n_spaxel, cube, (const long long*)spaxel_index, theta_deg, vel0, broad0, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_prior_kernel");
    return 0;
}

int lucifit_sanitize_cube(float* cube, int64_t n_values, float lo, float hi, float fill, void* stream)
{
    if (n_values < 0 || (!cube && n_values > 0)) return fail(LUCIFIT_EINVAL, "bad argument");
    if (!(lo <= hi)) return fail(LUCIFIT_EINVAL, "lo must not exceed hi");
    if (n_values == 0) return 0;
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_sanitize(cube, n_values, lo, hi, fill, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_sanitize_kernel");
    return 0;
}

int lucifit_pack_maps(const float* rows, int64_t n_rows, int row_floats, float* maps, uint32_t* maps_be, void* stream)
{
    if (n_rows < 0 || row_floats < 1 || row_floats > 4096) return fail(LUCIFIT_EINVAL, "bad argument");
    if (n_rows == 0) return 0;
    if (!rows || !maps) return fail(LUCIFIT_EINVAL, "rows and maps are required");
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_pack_maps(rows, n_rows, row_floats, maps, (unsigned*)maps_be, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_pack_maps_kernel");
    return 0;
}

int lucifit_row_median(const float* cube, int64_t n_rows, int nz, int lo, int hi, float* median, void* stream)
{
    if (n_rows < 0 || nz < 1 || lo < 0 || hi > nz || hi <= lo) return fail(LUCIFIT_EINVAL, "bad argument");
    if (hi - lo > 1024) return fail(LUCIFIT_EINVAL, "row_median: windows of up to 1024 samples are supported");
    if (n_rows == 0) return 0;
    if (!cube || !median) return fail(LUCIFIT_EINVAL, "cube and median are required");
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_rowmedian(cube, n_rows, nz, lo, hi, median, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_rowmedian_kernel");
    return 0;
}

int lucifit_segment_sum(const float* cube, int nz, const int64_t* row_index, const int32_t* segment, int64_t n_sel,
                        int n_segments, int nansum, double* sums, void* stream)
{
    if (nz < 1 || n_sel < 0 || n_segments < 1 || !sums) return fail(LUCIFIT_EINVAL, "bad argument");
    if (n_sel > 0 && !cube) return fail(LUCIFIT_EINVAL, "cube is required");
    int sm = 0; int rc = sm_count_of_current_device(&sm); if (rc) return rc;
    cudaError_t e = lf_run_segsum(cube, nz, (const long long*)row_index, (const int*)segment, n_sel, n_segments, nansum,
                                  sums, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_segsum_kernel");
    return 0;
}

// ---- peer memory (multi-GPU gather without a gather step) -----------------------------------------------------
// cudaIpcGetMemHandle names a whole allocation, and the caller's buffer usually lives inside a larger block of its
// framework's caching allocator: the block is found with cuMemGetAddressRange (fetched through the runtime's driver
// entry-point lookup, so the library does not link libcuda) and the buffer travels as (handle, offset).
typedef int (*lf_cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);
static lf_cuMemGetAddressRange_t lf_get_address_range_fn()
{
    static lf_cuMemGetAddressRange_t fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess && p)
            fn = (lf_cuMemGetAddressRange_t)p;
    }
    return fn;
}

int lucifit_peer_export(const void* dev_ptr, void* handle, uint64_t* offset)
{
    if (!dev_ptr || !handle || !offset) return fail(LUCIFIT_EINVAL, "dev_ptr, handle and offset are required");
    static_assert(sizeof(cudaIpcMemHandle_t) == LUCIFIT_PEER_HANDLE_BYTES, "handle size");
    lf_cuMemGetAddressRange_t range = lf_get_address_range_fn();
    if (!range) return fail(LUCIFIT_ECUDA, "cuMemGetAddressRange is not available");
    unsigned long long base = 0; size_t size = 0;
    if (range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0)
        return fail(LUCIFIT_ECUDA, "cuMemGetAddressRange failed (not a device allocation?)");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
    memcpy(handle, &h, sizeof(h));
    *offset = (uint64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return 0;
}

int lucifit_peer_open(const void* handle, uint64_t offset, void** dev_ptr)
{
    if (!handle || !dev_ptr) return fail(LUCIFIT_EINVAL, "handle and dev_ptr are required");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
    *dev_ptr = (char*)base + offset;
    {
        unsigned long long rb = 0; size_t size = 0;
        lf_cuMemGetAddressRange_t range = lf_get_address_range_fn();
        if (!range || range(&rb, &size, (unsigned long long)(uintptr_t)base) != 0) { rb = (unsigned long long)(uintptr_t)base; size = (size_t)1 << 46; }
        std::lock_guard<std::mutex> g(lf_peer_mu);
        lf_peer_ranges.push_back(std::make_pair((uintptr_t)rb, size));
    }
    return 0;
}

int lucifit_peer_close(void* dev_ptr, uint64_t offset)
{
    if (!dev_ptr) return 0;
    {
        std::lock_guard<std::mutex> g(lf_peer_mu);
        const uintptr_t b = (uintptr_t)((char*)dev_ptr - offset);
        for (size_t i = 0; i < lf_peer_ranges.size(); ++i)
            if (lf_peer_ranges[i].first == b) { lf_peer_ranges.erase(lf_peer_ranges.begin() + i); break; }
    }
    cudaError_t e = cudaIpcCloseMemHandle((char*)dev_ptr - offset);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
    return 0;
}

int lucifit_peak_probe(int mode, int iters, int blocks, float* out, void* stream)
{
    if ((mode != 0 && mode != 1) || iters < 1 || blocks < 1 || !out) return fail(LUCIFIT_EINVAL, "bad argument");
    cudaError_t e = lf_run_peak(mode, iters, blocks, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch lf_peak_kernel");
    return 0;
}

}  // extern "C"
