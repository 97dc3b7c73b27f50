/* lucifit.h -- C ABI of the B200-native LUCI fitting path (liblucifit.so).
 *
 * LUCI (crhea93/LUCI) is pure Python and has no FFI of its own; the seam this library plugs into is the
 * process boundary of its row worker:
 *
 *   joblib.Parallel(...)(delayed(Luci.fit_calc)(row, ...))        LuciBase.py:514-530, 691-705
 *   Luci.fit_calc(i, x_min, x_max, y_min, fit_function, lines, vel_rel, sigma_rel, cube_slice, ...)
 *       -> (i, ampls, flux, flux_errs, vels, vels_errs, broads, broads_errs, chi2, corr, step,
 *           continuum, continuum_errs)                              LuciBase.py:258-268, 406
 *   np.nansum(cube, axis=2) chunks (create_deep_image)              LuciBase.py:171-178
 *   SNR_calc(i) per row (create_snr_map)                            LuciBase.py:1109-1175
 *   bin_cube_function                                               LUCI/LuciUtility.py:319-355
 *
 * One lucifit_fit_batch call replaces the whole Parallel(...) loop for one GPU's shard of spaxels.
 *
 * Conventions: plain C, no exceptions cross the boundary; every function returns 0 on success or a negative
 * LUCIFIT_E* code and leaves a message for lucifit_last_error() (thread-local).  The library never allocates
 * or frees caller-visible memory: all buffers are DEVICE pointers owned by the caller (PyTorch tensors in the
 * Python glue), plus a caller-provided device workspace whose size is queried first.  All work is enqueued on
 * the caller's stream (passed as void* == cudaStream_t); calls are re-entrant per (device, stream).
 */
#ifndef LUCIFIT_H
#define LUCIFIT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUCIFIT_MAX_LINES 8

#define LUCIFIT_OK 0
#define LUCIFIT_EINVAL (-1)      /* bad argument / unsupported configuration */
#define LUCIFIT_EWORKSPACE (-2)  /* workspace too small */
#define LUCIFIT_ECUDA (-3)       /* a CUDA runtime call failed */

enum { LUCIFIT_GAUSSIAN = 0, LUCIFIT_SINC = 1, LUCIFIT_SINCGAUSS = 2 };

/* Per-spaxel status word written to out_status: bits 0-7 accepted LM iterations, bits 8-15 model+Jacobian
 * evaluations, flags from bit 16. */
#define LUCIFIT_ST_CONVERGED (1 << 16)
#define LUCIFIT_ST_MAXITER (1 << 17)
#define LUCIFIT_ST_BOUND (1 << 18)       /* a parameter sits on its box (LuciFit.py:614-633, 539-541) */
#define LUCIFIT_ST_SINGULAR (1 << 19)
#define LUCIFIT_ST_NAN_INPUT (1 << 20)   /* every sample is NaN: zeros are returned (LuciBase.py:365, 393-405) */
#define LUCIFIT_ST_MASKED (1 << 21)
#define LUCIFIT_ST_BAD_NORM (1 << 22)    /* max(spectrum) <= 0 */
#define LUCIFIT_ST_NAN_SAMPLES (1 << 24) /* informational: NaN samples were left out of the fit, the rest was fitted like the
                                            reference's row worker does after sky[~isnan(sky)] (LuciBase.py:358-360) */

/* Per-call constants (HOST memory).  Mirrors the arguments of Fit.__init__ (LUCI/LuciFit.py:49-57) that do not
 * vary from spaxel to spaxel. */
typedef struct lucifit_config {
    int model;                                /* LUCIFIT_GAUSSIAN / _SINC / _SINCGAUSS   (model_type) */
    int n_lines;                              /* L <= LUCIFIT_MAX_LINES                  (lines) */
    double line_nm[LUCIFIT_MAX_LINES];        /* rest wavelengths, nm                    (line_dict, LuciFit.py:90-99) */
    int vel_group[LUCIFIT_MAX_LINES];         /* vel_rel   (any integer labels) */
    int sigma_group[LUCIFIT_MAX_LINES];       /* sigma_rel (any integer labels) */
    int nii6548_idx, nii6583_idx;             /* indices of the [NII] doublet when nii_cons applies, else -1 */
    int nz;                                   /* spectral samples per spaxel */
    const double* axis;                       /* [nz] HOST: float32-rounded axis * (1 + obj_redshift), widened */
    int fit_lo, fit_hi;                       /* argmin|axis - spec_min|, argmin|axis - spec_max|  (LuciFit.py:273-274) */
    int noise_lo, noise_hi;                   /* noise window indices                              (LuciFit.py:327-328) */
    int cont_lo, cont_hi;                     /* continuum window indices                          (LuciFit.py:476-477) */
    double step_nm;                           /* hdr['STEP']   (delta_x) */
    double n_steps;                           /* hdr['STEPNB'] (n_steps) */
    double zpd_index;                         /* hdr['ZPDINDEX'] */
    const double* transmission;               /* [nz] HOST transmission_interpolated, or NULL (LuciFit.py:197-207) */
    int uncertainty;                          /* uncertainty_bool */
    int scale_mode;                           /* 1: ML-prior path scale (LuciFit.py:374-375); 0: max(spectrum) (:782) */
    int max_iter;                             /* LM iteration cap (0 -> 60) */
    double ftol;                              /* predicted relative decrease of sum r^2 that counts as converged (0 -> 1e-8) */
    int prior_groups;                         /* 0: vel0 / broad0 hold one value per spaxel (the reference's CNN output);
                                                 1: vel0 is [n_spaxel][NV], broad0 [n_spaxel][NS], one prior per tie group in
                                                 order of first appearance in vel_group / sigma_group (double components) */
    int freeze;                               /* 1: the reference's freeze branch (initial_values, LuciFit.py:159-162, 688-709):
                                                 vel0 / broad0 are FROZEN, only amplitudes and continuum are fitted, without
                                                 ties or boxes, continuum guess clipped at 3 sigma; sincgauss only (the
                                                 reference's gaussian / sinc freeze paths raise IndexError) */
    int no_prior_refine;                      /* 0 (default): the prep stage moves the prior velocity to the maximum of a matched
                                                 filter within +-36 km/s of it (the grid follows a maximum on its edge, up to +-108 km/s)
                                                 before the optimiser starts (fewer iterations, same optimum; skipped with
                                                 prior_groups / freeze / several velocity groups);
                                                 1: start exactly at vel0 */
} lucifit_config;

/* Number of floats per spaxel in out_rows: 7 L + 5, ordered
 *   amplitudes[L] fluxes[L] flux_errors[L] velocities[L] vels_errors[L] sigmas[L] sigmas_errors[L]
 *   chi2 corr axis_step continuum continuum_error           (the 12 lists of Luci.fit_calc, LuciBase.py:406) */
int lucifit_row_floats(int n_lines);

/* Device workspace needed by lucifit_fit_batch: constant tables plus the per-spaxel state records of one slice
 * (min(n_spaxel, 2^20) spaxels; larger batches are processed slice by slice). */
int lucifit_workspace_bytes(const lucifit_config* cfg, int64_t n_spaxel, size_t* bytes);

/* Fit n_spaxel spectra.
 *   cube         [*, nz] float32 DEVICE, z contiguous; spaxel s reads row (spaxel_index ? spaxel_index[s] : s)
 *   spaxel_index optional DEVICE int64 list (mask compaction) or NULL
 *   theta_deg    [n_spaxel] DEVICE interferometer angle of each fitted spaxel (LuciBase.py:369)
 *   vel0, broad0 [n_spaxel] DEVICE prior velocity / broadening in km/s (what the CNN writes to vel_ml / broad_ml,
 *                LuciFit.py:356-361), either may be NULL (-> 0, the reference's ML_bool=False behaviour); with
 *                cfg->prior_groups they are [n_spaxel][NV] / [n_spaxel][NS]
 *   bkg          [nz] DEVICE background spectrum already multiplied by binning^2, or NULL (LuciBase.py:326-330)
 *   out_rows     [n_spaxel][7L+5] DEVICE float32
 *   out_fit_sol  optional [n_spaxel][3L+1] DEVICE: fit_sol (amplitude, position cm^-1, sigma cm^-1)*L, continuum
 *   out_unc      optional [n_spaxel][3L+1] DEVICE: fit_uncertainties in the same layout
 *   out_status   [n_spaxel] DEVICE int32 status words
 */
int lucifit_fit_batch(const lucifit_config* cfg, int64_t n_spaxel, const float* cube,
                      const int64_t* spaxel_index, const float* theta_deg, const float* vel0, const float* broad0,
                      const float* bkg, float* out_rows, float* out_fit_sol, float* out_unc, int32_t* out_status,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The same pipeline run stage by stage (each stage is one kernel launch; a spaxel travels between stages as a
 * state record inside `workspace`, so the calls of one pipeline must use the same cfg / workspace / n_spaxel, in
 * order, on one stream; n_spaxel <= 2^20 when `stages` is not LUCIFIT_STAGE_ALL).  lucifit_fit_batch is
 * lucifit_fit_stages(cfg, LUCIFIT_STAGE_ALL, ...).  Used to time or profile one stage (bench.py) and to interleave
 * other work between stages.
 *   PREP   Fit.__init__ + cont_estimate + line_vals_estimate   (LuciFit.py:49-173, 382-486)
 *   LM     calculate_params, the optimiser                      (LuciFit.py:635-687)
 *   UNC    uncertainties (only when cfg->uncertainty)           (LuciFit.py:713-722, LuciUtility.py:409-434)
 *   OUT    chi2, fluxes, velocities, broadenings, rows          (LuciFit.py:799-834, LuciFitParameters.py) */
#define LUCIFIT_STAGE_PREP 1
#define LUCIFIT_STAGE_LM 2
#define LUCIFIT_STAGE_UNC 4
#define LUCIFIT_STAGE_OUT 8
#define LUCIFIT_STAGE_ALL 15
int lucifit_fit_stages(const lucifit_config* cfg, int stages, int64_t n_spaxel, const float* cube,
                       const int64_t* spaxel_index, const float* theta_deg, const float* vel0, const float* broad0,
                       const float* bkg, float* out_rows, float* out_fit_sol, float* out_unc, int32_t* out_status,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Initial velocity / broadening (km/s) of n_spaxel spectra, measured from the data: what the reference obtains from
 * its Keras CNN (Fit.estimate_priors_ML, LUCI/LuciFit.py:335-362, consumed by line_vals_estimate :382-430) and feeds
 * to the optimiser as vel_ml / broad_ml.  Velocity: matched filter of the continuum-subtracted fit window against
 * the expected centres of cfg's lines over the grid v_min, v_min + v_step, ... <= v_max (at most 512 points),
 * refined by a parabola; broadening: FWHM of the strongest line with the instrumental sinc width removed.
 * Uses cfg->model, line_nm, axis, fit_lo / fit_hi, step_nm, n_steps, zpd_index.  cube / spaxel_index / theta_deg / bkg
 * as in lucifit_fit_batch; vel0, broad0: [n_spaxel] DEVICE outputs, directly usable as its vel0 / broad0. */
int lucifit_estimate_prior(const lucifit_config* cfg, int64_t n_spaxel, const float* cube, const int64_t* spaxel_index,
                           const float* theta_deg, const float* bkg, double v_min, double v_max, double v_step,
                           float* vel0, float* broad0, void* stream);

/* Deep image: deep[s] = nansum_z cube[s][z] for s < n_spaxel - n_zero_tail, 0 for the rest
 * (create_deep_image leaves the trailing nx mod 10 x-rows zero, LuciBase.py:173-177). */
int lucifit_deep_image(const float* cube, int64_t n_spaxel, int nz, int64_t n_zero_tail, float* deep,
                       void* stream);

/* SNR map (LuciBase.py:1142-1172).  method 1: (nanmax - nanmean)/sqrt(std(noise window)), 0 if negative, else
 * divided by sqrt(nanmean|sky|).  method 2: max(0, (nansum(flux window) - n * min(sigma_clip(sky, 1, 10 iters)))
 * / std(noise window)).  If deep != NULL the same pass also writes the deep image (cube read once). */
int lucifit_snr_map(const float* cube, int64_t n_spaxel, int nz, int method, int flux_lo, int flux_hi,
                    int noise_lo, int noise_hi, float* snr, float* deep, void* stream);

/* b x b nansum binning of the region [x0,x1) x [y0,y1) of cube[nx][ny][nz] -> binned[(x1-x0)/b][(y1-y0)/b][nz]
 * (bin_cube_function, LUCI/LuciUtility.py:332-341; no division). */
int lucifit_bin_cube(const float* cube, int nx, int ny, int nz, int b, int x0, int x1, int y0, int y1,
                     float* binned, void* stream);

/* Cube ingest, in place: the clamp rules Luci.read_in_cube applies to every HDF5 quadrant (LuciBase.py:123-129) --
 * non-finite samples, samples < lo (-1e-16) and samples > hi (1e-9) become `fill` (1e-22).  cube: n_values DEVICE floats. */
int lucifit_sanitize_cube(float* cube, int64_t n_values, float lo, float hi, float fill, void* stream);

/* Synthetic LuciSim-style cube written straight into HBM (benchmark input; LUCI/LuciSim.py:159-203):
 * spectrum[s][z] = flux_scale * (continuum + sum_k amp[s][k] * profile(axis[z]; centre snapped to the axis,
 * sigma from broad[s][k]) + noise_sigma[s] * noise[s][z]) where noise is caller-provided N(0,1) draws (or NULL;
 * it may alias cube: every element is read and then written by the same thread).  vel/broad follow LuciSim's
 * conventions: centre = (vel/3e5 + 1) 1e7/lambda, sigma = broad * centre / (3e5 corr). */
int lucifit_sim_cube(const lucifit_config* cfg, int64_t n_spaxel, const float* amp /*[n][L]*/,
                     const float* vel /*[n][L]*/, const float* broad /*[n][L]*/, const float* theta_deg /*[n]*/,
                     const float* noise_sigma /*[n]*/, const float* noise /*[n][nz] or NULL*/, float continuum,
                     float flux_scale, float* cube /*[n][nz]*/, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Output maps: rows[n_rows][row_floats] (the out_rows of lucifit_fit_batch, one row per spaxel in map order) ->
 * maps[row_floats][n_rows], the field-major layout of the 2-D maps the reference scatters its workers' lists into and
 * hands to save_fits (LuciBase.py:531-551, LUCI/LuciUtility.py:69-85).  maps_be (optional) receives the same values with
 * the bytes of every float reversed (big-endian, the byte order of a FITS data unit), so that the files can be written
 * from it without touching the data on the host.  All DEVICE pointers. */
int lucifit_pack_maps(const float* rows, int64_t n_rows, int row_floats, float* maps, uint32_t* maps_be, void* stream);

/* median[r] = np.nanmedian(cube[r][lo:hi]) (the two middle values averaged in float64; NaN when nothing is valid): the
 * level the absorption template is scaled by, LuciBase.py:354-356.  hi - lo <= 1024.  DEVICE pointers. */
int lucifit_row_median(const float* cube, int64_t n_rows, int nz, int lo, int hi, float* median, void* stream);

/* Integrated spectra: sums[segment[i]][z] += cube[row_index[i]][z] for i < n_sel, float64 -- the spaxel loops of
 * fit_spectrum_region (LuciBase.py:1015-1023), extract_spectrum / extract_spectrum_region (:872-889, :931-941), the
 * per-bin spectra of fit_wvt (:1434-1447) and the region spectra of skyline_calibration (:1254-1256) as ONE pass.
 * row_index: [n_sel] DEVICE rows of cube[.][nz] (NULL: rows 0..n_sel-1); segment: [n_sel] DEVICE ids in
 * [0, n_segments), negative = skip (NULL: everything into segment 0); a list ordered by segment is fastest.
 * nansum != 0 skips NaN samples (np.nansum), 0 lets them poison the channel like `+=` in the reference.
 * sums: [n_segments][nz] DEVICE doubles, zeroed by the call. */
int lucifit_segment_sum(const float* cube, int nz, const int64_t* row_index, const int32_t* segment, int64_t n_sel,
                        int n_segments, int nansum, double* sums, void* stream);

/* Peer rows -- the multi-GPU "gather" of a sharded fit without a gather step.  The reference collects the rows of its
 * joblib workers on the host (LuciBase.py:514-551); here every rank's lucifit_fit_batch writes its (7L+5)-float rows
 * and status words STRAIGHT into the destination rank's map buffer over NVLink: the destination exports its buffer,
 * the other processes open it and pass the mapped pointer (plus their row offset) as out_rows / status.
 *   lucifit_peer_export: (handle, offset) naming dev_ptr for other processes of this node; `handle` receives
 *                        LUCIFIT_PEER_HANDLE_BYTES bytes (a cudaIpcMemHandle_t of the enclosing allocation).
 *   lucifit_peer_open:   maps it into this process on the CURRENT device and returns the pointer.
 *   lucifit_peer_close:  unmaps (same pointer and offset as returned / given to lucifit_peer_open).
 * Completion: a rank's rows are in place when its stream has drained; the ranks then meet at a barrier.
 * The library keeps the ranges it has mapped: when out_rows of lucifit_fit_batch / lucifit_fit_stages lies inside one,
 * the output kernel assembles each row in shared memory and writes it as whole 128-byte lines (2 NVLink writes per row
 * instead of 12 small ones). */
#define LUCIFIT_PEER_HANDLE_BYTES 64
int lucifit_peer_export(const void* dev_ptr, void* handle, uint64_t* offset);
int lucifit_peer_open(const void* handle, uint64_t offset, void** dev_ptr);
int lucifit_peer_close(void* dev_ptr, uint64_t offset);

/* Roofline probe: `blocks` CTAs of 256 threads each run `iters` iterations of 8 independent FFMA (mode 0) or
 * MUFU.EX2 (mode 1) chains; out holds blocks*256 floats.  Work = blocks*256*iters*8 FMA (x2 flop) or MUFU ops.
 * MEASURED_PEAKS.json has no FP32 / SFU figure; bench.py times this to get the denominators of the fit kernel. */
int lucifit_peak_probe(int mode, int iters, int blocks, float* out, void* stream);

const char* lucifit_last_error(void);
const char* lucifit_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LUCIFIT_H */
