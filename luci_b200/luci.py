"""``Luci`` -- drop-in for the fitting / reduction methods of ``LuciBase.Luci`` (reference ``LuciBase.py:44-1191``).

Same method names, argument meaning and return values as the reference for the hot path
(``fit_entire_cube``, ``fit_cube``, ``fit_region``, ``fit_spectrum_region``, ``create_deep_image``,
``create_snr_map``, ``bin_cube``), same FITS/NPY artefacts on disk (``LuciUtility.save_fits`` layout).  The joblib
fan-out over y-rows (``LuciBase.py:514-530``) is replaced by one ``lucifit_fit_batch`` call on the GPU holding the
cube.  The cube lives in HBM as float32 ``[x, y, z]`` (z contiguous), exactly the reference's ``cube_final`` axes.

Differences that are deliberate and documented in DESIGN.md:
  * ``Luci.from_arrays(...)`` builds the object from in-memory arrays (h5py/astropy are optional here).
  * the CNN prior of the reference (TensorFlow, out of scope) enters as ``prior_values=(vel_map, broad_map)`` in
    km/s, which is exactly what the CNN writes into ``vel_ml`` / ``broad_ml`` (``LuciFit.py:352-361``); without
    ``prior_values`` and with ``ML_bool=True`` the two numbers are measured from each spectrum on the device
    (``engine.estimate_prior`` -> ``lucifit_estimate_prior``: matched-filter velocity, line-width broadening);
  * ``fit_entire_cube`` honours its keyword arguments and returns the ``fit_cube`` tuple (the reference drops them
    and returns ``None``, ``LuciBase.py:255``);
  * the Bayesian branch (``bayes_bool``) raises ``NotImplementedError`` (out of scope); PCA background, absorption
    template, freeze mode (``initial_values``) and ``n_stoch`` restarts are implemented for the cube fits.
"""
import os
import warnings

import numpy as np
import numpy.ma as ma
import torch

from . import dist as ldist
from . import engine
from ._capi import FitConfig, argmin_abs
from .constants import LINE_DICT, snr_windows
from .fitsio import read_fits, save_fits, wcs_to_header, write_fits


def _nvtx(name):
    """Decorator: an NVTX range around a method (visible in Nsight captures next to the library's own lucifit:prep / lm /
    unc / out ranges; the reference has only tqdm bars, LuciBase.py:530)."""
    def wrap(fn):
        import functools

        @functools.wraps(fn)
        def inner(*a, **k):
            on = torch.cuda.is_available()
            if on:
                torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                if on:
                    torch.cuda.nvtx.range_pop()
        return inner
    return wrap


SKYLINE_V0 = 80.0          # km/s: the reference's initial guess for the sky-line velocity (LuciFit.py:859)
# Component search of a multi-component fit started from ONE prior pair (all the reference's CNN can give): the same
# line in two velocity groups starts exactly symmetric, a saddle of the objective the reference only leaves through the
# noise of its finite-difference gradient ("the results are not always correct", docs/source/example_double_fit.rst).
# Here every further component is started at the prior plus each of these offsets and the lowest chi2 wins per spaxel.
COMPONENT_OFFSETS = (-300.0, -200.0, -100.0, 100.0, 200.0, 300.0)


def check_luci_path(Luci_path):
    if Luci_path is not None and not Luci_path.endswith('/'):
        Luci_path += '/'
        print("We have added a trailing '/' to your Luci_path variable.\n")
        print("Please add this in the future.\n")
    return Luci_path


def spectrum_axis_func(hdr_dict, redshift):
    """``LUCI/LuciUtility.py:143-165``: float32 axis (redshifted and unshifted) from the header."""
    len_wl = int(hdr_dict['STEPNB'])
    start = float(hdr_dict['CRVAL3'])
    step = float(hdr_dict['CDELT3'])
    end = start + len_wl * step
    axis = np.array(np.linspace(start, end, len_wl) * (redshift + 1), dtype=np.float32)
    axis_unshifted = np.array(np.linspace(start, end, len_wl), dtype=np.float32)
    return axis, axis_unshifted


class Luci():
    STAGE_MIN_VALUES = 1 << 22      # maps with at least this many values travel through the pinned staging buffers

    def __init__(self, Luci_path, cube_path, output_dir, object_name, redshift, resolution, ML_bool=True, mdn=False,
                 device=None):
        """Reference signature (``LuciBase.py:51``): reads ``<cube_path>.hdf5`` (needs ``h5py``, see
        :mod:`luci_b200.hdf5io`); use :meth:`from_arrays` for in-memory data.  The reference spectrum of the CNN
        (``wavenumbers_syn``, ``LuciBase.py:98-99``) is not read: the CNN prior is replaced (see ``prior_values``)."""
        self._init_common(Luci_path, output_dir, object_name, redshift, resolution, ML_bool, mdn, device)
        self.cube_path = cube_path
        self.read_in_cube()
        self._finish_init()

    # ------------------------------------------------------------------------------------------ construction
    def _init_common(self, Luci_path, output_dir, object_name, redshift, resolution, ML_bool, mdn, device):
        self.header_binned = None
        self.Luci_path = check_luci_path(Luci_path)
        self.cube_path = None
        self.output_dir = output_dir + '/Luci_outputs'
        os.makedirs(self.output_dir, exist_ok=True)
        self.object_name = object_name
        self.redshift = redshift
        self.resolution = resolution
        self.mdn = mdn
        self.ML_bool = ML_bool
        self.cube_final = None          # torch float32 [x, y, z] on the device
        self.cube_binned = None
        self.header = None              # dict of FITS cards (WCS)
        self.deep_image = None
        self.spectrum_axis = None
        self.spectrum_axis_unshifted = None
        self.wavenumbers_syn = None
        self.hdr_dict = None
        self.interferometer_theta = None
        self.transmission_interpolated = None
        self.strict_reference = True    # replicate the reference's quirks that change numbers
        self.last_fit_summary = None
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.device = torch.device(device) if device is not None else None

    def _finish_init(self):
        self.step_nb = self.hdr_dict['STEPNB']
        self.zpd_index = self.hdr_dict['ZPDINDEX']
        self.filter = self.hdr_dict['FILTER']
        if self.spectrum_axis is None:
            self.spectrum_axis, self.spectrum_axis_unshifted = spectrum_axis_func(self.hdr_dict, self.redshift)
            if self.filter in ('C4', 'C2', 'C1'):
                self.spectrum_axis = self.spectrum_axis_unshifted      # LuciBase.py:96-97
        self.dimx, self.dimy, self.dimz = (int(v) for v in self.cube_final.shape)

    @classmethod
    def from_arrays(cls, cube, hdr_dict, output_dir, object_name, redshift=0.0, resolution=5000,
                    interferometer_theta=None, spectrum_axis=None, transmission=None, header=None, Luci_path=None,
                    ML_bool=True, mdn=False, device=None):
        """Build a ``Luci`` from in-memory data.

        cube: ``[x, y, z]`` numpy array or torch tensor (copied to the GPU as float32; a CUDA float32 tensor is
        used in place).  hdr_dict needs ``FILTER``, ``STEP``, ``STEPNB``, ``ZPDINDEX`` and, unless
        ``spectrum_axis`` is given, ``CRVAL3``/``CDELT3``.  interferometer_theta: ``[x, y]`` degrees
        (default 11.96, the LuciSim angle).  transmission: ``[z]`` or ``None``.
        """
        self = cls.__new__(cls)
        self._init_common(Luci_path, output_dir, object_name, redshift, resolution, ML_bool, mdn, device)
        if self.device is None:
            raise RuntimeError('luci_b200 needs a CUDA device: there is no CPU path')
        self.hdr_dict = dict(hdr_dict)
        self.header = dict(header) if header is not None else {}
        self.cube_final = self._to_device_cube(cube)
        nx, ny = self.cube_final.shape[:2]
        if interferometer_theta is None:
            interferometer_theta = np.full((nx, ny), 11.96)
        self.interferometer_theta = np.asarray(interferometer_theta, dtype=np.float64)
        if spectrum_axis is not None:
            self.spectrum_axis = np.asarray(spectrum_axis, dtype=np.float32)
            self.spectrum_axis_unshifted = self.spectrum_axis
        self.transmission_interpolated = None if transmission is None else np.asarray(transmission, dtype=np.float64)
        self._finish_init()
        return self

    def _to_device_cube(self, cube):
        if isinstance(cube, torch.Tensor):
            t = cube
            if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
                return t
            return t.to(device=self.device, dtype=torch.float32).contiguous()
        cube = np.asarray(cube)
        out = torch.empty(cube.shape, dtype=torch.float32, device=self.device)
        step = max(1, int(2 ** 28 // max(1, cube.shape[1] * cube.shape[2])))      # ~1 GiB float32 chunks along x
        for i in range(0, cube.shape[0], step):
            out[i:i + step] = torch.from_numpy(np.ascontiguousarray(cube[i:i + step], dtype=np.float32)).to(self.device)
        return out

    def read_in_cube(self):
        """``LuciBase.py:106-145``: HDF5 quadrants -> ``cube_final[x, y, z]`` (float32, in HBM, clamp rules of
        :123-129 applied by one device pass), header -> ``hdr_dict`` / WCS ``header``, calibration map ->
        ``interferometer_theta``; then the transmission curve of the filter (``:100-101``).  Needs ``h5py``."""
        from . import hdf5io
        if self.device is None:
            raise RuntimeError('luci_b200 needs a CUDA device: there is no CPU path')
        print('Reading in data...')
        file = hdf5io.open_hdf5(self.cube_path)
        self.cube_final = hdf5io.read_cube(file, self.device)
        self.hdr_dict = hdf5io.header_dict(file)
        self.header = hdf5io.wcs_header(self.hdr_dict)
        self.interferometer_theta = hdf5io.interferometer_angles(file, self.hdr_dict)
        self.spectrum_axis, self.spectrum_axis_unshifted = spectrum_axis_func(self.hdr_dict, self.redshift)
        if self.hdr_dict['FILTER'] in ('C4', 'C2', 'C1'):
            self.spectrum_axis = self.spectrum_axis_unshifted          # LuciBase.py:96-97
        if self.Luci_path is not None:
            self.transmission_interpolated = hdf5io.read_in_transmission(self.Luci_path, self.hdr_dict,
                                                                         self.spectrum_axis_unshifted)

    # -------------------------------------------------------------------------------------------- reductions
    @_nvtx('luci:create_deep_image')
    def create_deep_image(self, output_name=None, binning=None):
        """``LuciBase.py:147-215``: sum over the spectral axis, transposed to (ny, nx), written as FITS."""
        nx, ny, nz = self.cube_final.shape
        tail = (nx % 10) * ny if self.strict_reference else 0          # reference leaves nx mod 10 rows zero
        deep = engine.deep_image(self.cube_final.view(nx * ny, nz), n_zero_tail=tail)
        self.deep_image = deep.view(nx, ny).t().contiguous().cpu().numpy().astype(np.float64)
        header_to_use = self.header
        if binning is not None and binning != 1:
            b = int(binning)
            xs, ys = int(self.deep_image.shape[0] / b), int(self.deep_image.shape[1] / b)
            d = self.deep_image[:xs * b, :ys * b].reshape(xs, b, ys, b)
            self.deep_image = np.nansum(np.nansum(d, axis=3), axis=1) / (b ** 2)
            header_to_use = dict(self.header)
            for key, f in (('CRPIX1', 1.0 / b), ('CRPIX2', 1.0 / b), ('PC1_1', b), ('PC1_2', b), ('PC2_1', b), ('PC2_2', b)):
                if key in header_to_use:
                    header_to_use[key] = header_to_use[key] * f
        if output_name is None:
            output_name = self.output_dir + '/' + self.object_name + '_deep.fits'
        write_fits(output_name, self.deep_image, wcs_to_header(header_to_use))

    def bin_cube(self, cube_final, header, binning, x_min, x_max, y_min, y_max):
        """``LuciBase.py:840-842`` / ``LuciUtility.py:319-355``: b x b nansum (no division)."""
        self.cube_binned = engine.bin_cube(cube_final, binning, x_min, x_max, y_min, y_max)
        hb = dict(header) if header else {}
        if 'CRPIX1' in hb:
            hb['CRPIX1'] = (hb['CRPIX1'] - x_min - 0.5) / binning + 0.5
        if 'CRPIX2' in hb:
            hb['CRPIX2'] = (hb['CRPIX2'] - y_min - 0.5) / binning + 0.5
        for key in ('PC1_1', 'PC1_2', 'PC2_1', 'PC2_2'):
            if key in hb:
                hb[key] = hb[key] * binning
        self.header_binned = hb        # the reference mutates self.header in place here; that bug is not replicated

    @_nvtx('luci:create_snr_map')
    def create_snr_map(self, x_min=0, x_max=2048, y_min=0, y_max=2064, method=1, n_threads=2, lines=[None],
                       binning=1, bkgType=None, pca_coefficient_array=None, pca_vectors=None, pca_mean=None):
        """``LuciBase.py:1048-1191``.  Returns the four masked arrays (SNR >= 1, 3, 5, 10) like the reference."""
        cube_to_use = self.cube_final
        if binning > 1:
            self.bin_cube(self.cube_final, self.header, binning, x_min, x_max, y_min, y_max)
            x_max = int((x_max - x_min) / binning)
            y_max = int((y_max - y_min) / binning)
            x_min = 0
            y_min = 0
            cube_to_use = self.cube_binned
        (fmin, fmax), (nmin, nmax) = snr_windows(self.hdr_dict['FILTER'], lines)
        ax = np.array(self.spectrum_axis)
        flo, fhi = argmin_abs(ax, fmin), argmin_abs(ax, fmax)
        nlo, nhi = argmin_abs(ax, nmin), argmin_abs(ax, nmax)
        x_max = min(x_max, cube_to_use.shape[0]); y_max = min(y_max, cube_to_use.shape[1])
        region = cube_to_use[x_min:x_max, y_min:y_max]
        nxr, nyr, nz = region.shape
        region2d = region.reshape(nxr * nyr, nz) if region.is_contiguous() else region.contiguous().view(nxr * nyr, nz)
        snr = engine.snr_map(region2d, method, flo, fhi, nlo, nhi)
        SNR = snr.view(nxr, nyr).t().contiguous().cpu().numpy()
        os.makedirs(self.output_dir + '/SNR', exist_ok=True)
        write_fits(self.output_dir + '/SNR/' + self.object_name + '_SNR.fits', SNR, wcs_to_header(self.header))
        masks = []
        for snr_val in [1, 3, 5, 10]:
            mask = ma.masked_where(SNR >= snr_val, SNR)
            masks.append(mask)
            np.save("%s/SNR/%s_SNR_%i_mask.npy" % (self.output_dir, self.object_name, snr_val), ma.getmaskarray(mask))
        self.snr_map = SNR
        return masks

    # --------------------------------------------------------------------------------------------------- fits
    def _make_config(self, lines, fit_function, vel_rel, sigma_rel, uncertainty_bool, nii_cons, spec_min, spec_max,
                     obj_redshift, ml_scale, prior_groups=False, freeze=False):
        return FitConfig(fit_function, lines, vel_rel, sigma_rel, self.spectrum_axis, filter=self.hdr_dict['FILTER'],
                         delta_x=self.hdr_dict['STEP'], n_steps=self.step_nb, zpd_index=self.zpd_index,
                         trans_filter=self.transmission_interpolated, uncertainty_bool=uncertainty_bool,
                         nii_cons=nii_cons, spec_min=spec_min, spec_max=spec_max, obj_redshift=obj_redshift,
                         ml_scale=ml_scale, prior_groups=prior_groups, freeze=freeze)

    @staticmethod
    def _reject_unsupported(bayes_bool, absorp, initial_values, n_stoch, bkgType):
        if bayes_bool:
            raise NotImplementedError('bayes_bool=True (emcee/dynesty MCMC, LuciFit.py:930-1015) is out of scope')
        if int(n_stoch) < 1:
            raise ValueError('n_stoch must be >= 1')

    def _cutout_header(self, header, x_min, y_min, x_max=None, y_max=None, full_shape=None):
        """The header ``Cutout2D(deep, position, size, wcs=WCS(header)).wcs.to_header()`` gives the reference's maps
        (``LuciBase.py:511-513``): WCSLIB's canonical card set (``fitsio.wcs_to_header``) with the reference pixel moved
        to the cut-out.  ``strict_reference`` reproduces ``Cutout2D``'s reading of ``size`` as ``(ny, nx)``, which
        displaces the origin of a non-square region (``fitsio.cutout_origin``); otherwise the fitted rectangle's own
        origin is used."""
        from .fitsio import cutout_origin
        if x_max is None or y_max is None:
            origin = (float(x_min), float(y_min))
        else:
            nx_f, ny_f = full_shape if full_shape is not None else self.cube_final.shape[:2]
            origin = cutout_origin(x_min, x_max, y_min, y_max, nx_f, ny_f, strict_reference=self.strict_reference)
        return wcs_to_header(header, origin)

    @_nvtx('luci:fit_maps')
    def _fit_maps(self, cube3d, lines, fit_function, vel_rel, sigma_rel, x_min, x_max, y_min, y_max, mask, bkg,
                  bkg_scale, uncertainty_bool, nii_cons, spec_min, spec_max, obj_redshift, prior_values, n_stoch=1,
                  pca=None, initial_values=(False,), absorp=None):
        """Shared body of fit_cube / fit_region: returns the 12 maps as numpy float32 arrays (ny', nx'[, L]).

        ``pca = (coefficients [n_fit_x, n_fit_y, n_comp] already summed over the bin, vectors [n_comp, nz], mean [nz])``
        selects the reference's PCA background (LuciBase.py:331-352)."""
        L = len(lines)
        nxc, nyc, nz = cube3d.shape
        nxo, nyo = x_max - x_min, y_max - y_min
        # initial_values = [velocity map, broadening map] (ny', nx') selects the reference's freeze branch
        # (LuciFit.py:159-162, 688-709; LuciBase.py:361-364): those values are held, amplitudes + continuum are fitted
        freeze = len(initial_values) >= 1 and initial_values[0] is not False
        if freeze:
            if uncertainty_bool:
                raise NotImplementedError('uncertainty_bool with initial_values is broken in the reference '
                                          '(LuciFit.py:713-722 on an L+1 vector) and not offered')
            prior_values = (np.asarray(initial_values[0], dtype=np.float32), np.asarray(initial_values[1], dtype=np.float32))
            n_stoch = 1
        have_prior = prior_values is not None
        # ML_bool=True without explicit priors: the reference asks its CNN (LuciFit.py:776-780); here the same two numbers
        # per spaxel are measured on the device (engine.estimate_prior).  ML_bool=False starts at velocity =
        # broadening = 0 like the reference, which only works for fit_function='sinc' (SURVEY fact 0.4).
        auto_prior = (not have_prior) and bool(self.ML_bool)
        if not have_prior and not auto_prior and fit_function != 'sinc':
            warnings.warn('ML_bool=False and no prior_values: every fit starts from velocity = broadening = 0 like the '
                          "reference, which only works for fit_function='sinc'")
        # priors with a trailing axis carry one value per tie group (order of first appearance in vel_rel /
        # sigma_rel): what a double-component fit needs to start its components apart
        per_group = have_prior and np.ndim(prior_values[0]) == 3
        # the same line in several velocity groups with one shared prior: component search (COMPONENT_OFFSETS)
        groups_of = {}
        for ln, vg in zip(lines, vel_rel):
            groups_of.setdefault(ln, set()).add(int(vg))
        search = (not per_group) and (not freeze) and any(len(v) > 1 for v in groups_of.values())
        if search:
            per_group = True
        cfg = self._make_config(lines, fit_function, vel_rel, sigma_rel, uncertainty_bool, nii_cons, spec_min,
                                spec_max, obj_redshift, ml_scale=have_prior or auto_prior, prior_groups=per_group,
                                freeze=freeze)
        if per_group and not search and (np.shape(prior_values[0])[2] != cfg.n_vel_groups or np.ndim(prior_values[1]) != 3
                                         or np.shape(prior_values[1])[2] != cfg.n_sigma_groups):
            raise Exception('per-group prior_values must be (ny, nx, %d) velocities and (ny, nx, %d) broadenings'
                            % (cfg.n_vel_groups, cfg.n_sigma_groups))
        dev = cube3d.device
        # fitted spaxels, ordered like the output maps: (y, x) row-major
        dense = mask is None and pca is None                # every spaxel of the rectangle: no selection arrays needed
        v0 = b0 = None
        pv = pb = None                                      # host priors of the fitted spaxels, [n_fit, 1 or groups]
        width = lambda a: 1 if np.ndim(a) < 3 else int(np.shape(a)[2])
        if dense:
            sel = None
            n_fit = nxo * nyo
            theta = np.ascontiguousarray(np.asarray(self.interferometer_theta)[x_min:x_max, y_min:y_max].T,
                                         dtype=np.float32).reshape(-1)              # LuciBase.py:369
            if have_prior:
                pv = np.ascontiguousarray(prior_values[0], dtype=np.float32).reshape(n_fit, width(prior_values[0]))
                pb = np.ascontiguousarray(prior_values[1], dtype=np.float32).reshape(n_fit, width(prior_values[1]))
        else:
            yy, xx = np.meshgrid(np.arange(y_min, y_max), np.arange(x_min, x_max), indexing='ij')
            if mask is not None:
                sel = np.asarray(mask)[xx, yy].astype(bool)
            else:
                sel = np.ones_like(xx, dtype=bool)
            xs_sel, ys_sel = xx[sel], yy[sel]
            n_fit = int(xs_sel.size)
            flat = xs_sel.astype(np.int64) * nyc + ys_sel.astype(np.int64)
            theta = np.asarray(self.interferometer_theta)[xs_sel, ys_sel].astype(np.float32)   # LuciBase.py:369
            if have_prior:
                pv = np.ascontiguousarray(np.asarray(prior_values[0], dtype=np.float32)[sel]).reshape(n_fit, width(prior_values[0]))
                pb = np.ascontiguousarray(np.asarray(prior_values[1], dtype=np.float32)[sel]).reshape(n_fit, width(prior_values[1]))
        # one process per GPU (torchrun): this rank fits a contiguous run [lo, hi) of the fitted list -- whole y-rows of
        # the output when the rectangle is dense, like the y-rows the reference hands to its joblib workers
        # (LuciBase.py:514-517), an even share of the compacted list under a mask -- and the rows meet on rank 0
        sharded = ldist.is_distributed()
        lo, hi = 0, n_fit
        if sharded:
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
            if dense:
                y_lo, y_hi = ldist.shard_bounds(nyo, rank, world)
                lo, hi = y_lo * nxo, y_hi * nxo
            else:
                lo, hi = ldist.split_even(n_fit, rank, world)
        if dense:
            y0, y1 = y_min + lo // nxo, y_min + hi // nxo
            index = (torch.arange(x_min, x_max, dtype=torch.int64, device=dev)[None, :] * nyc
                     + torch.arange(y0, y1, dtype=torch.int64, device=dev)[:, None]).reshape(-1)
        else:
            xs_sel, ys_sel = xs_sel[lo:hi], ys_sel[lo:hi]
            index = torch.from_numpy(flat[lo:hi]).to(dev)
        if have_prior:
            v0 = torch.from_numpy(np.ascontiguousarray(pv[lo:hi])).to(dev).reshape(-1)
            b0 = torch.from_numpy(np.ascontiguousarray(pb[lo:hi])).to(dev).reshape(-1)
        theta_t = torch.from_numpy(np.ascontiguousarray(theta[lo:hi])).to(dev)
        bkg_t = None
        if bkg is not None:
            bkg_t = torch.from_numpy(np.asarray(bkg, dtype=np.float32) * np.float32(bkg_scale)).to(dev)
        cube2d = cube3d.view(nxc * nyc, nz)
        if pca is not None:
            # PCA background (LuciBase.py:331-352): bkg = mean + coefficients . vectors per spaxel -- the one dense
            # contraction of the pipeline, a [n x n_comp] . [n_comp x nz] library GEMM -- scaled by the maximum of the
            # spectrum in the filter's reference band, subtracted into a compacted copy of the fitted spectra
            coef, vecs, mean = pca
            coef_t = torch.from_numpy(np.ascontiguousarray(np.asarray(coef, dtype=np.float32)[xs_sel - x_min, ys_sel - y_min])).to(dev)
            vecs_t = torch.from_numpy(np.ascontiguousarray(vecs, dtype=np.float32)).to(dev)
            mean_t = torch.from_numpy(np.ascontiguousarray(mean, dtype=np.float32)).to(dev)
            band = {'SN3': (675.0, 670.0), 'SN2': (505.0, 480.0)}.get(self.hdr_dict['FILTER'])
            if band is None:
                raise Exception('the PCA background is implemented for SN3 and SN2 (LuciBase.py:332-341)')
            nm_axis = 1e7 / np.asarray(self.spectrum_axis, dtype=np.float64)
            lo_i, hi_i = int(np.argmin(np.abs(nm_axis - band[0]))), int(np.argmin(np.abs(nm_axis - band[1])))
            sky = cube2d[index]
            bkg_pix = torch.addmm(mean_t[None, :].expand(sky.shape[0], nz), coef_t, vecs_t)
            ref_band = sky[:, lo_i:hi_i]
            scale_spec = torch.where(torch.isnan(ref_band), torch.full_like(ref_band, -float('inf')), ref_band).amax(dim=1)
            cube2d = (sky - scale_spec[:, None] * bkg_pix).contiguous()
            index = None
        if absorp is not None:
            # absorption template (LuciBase.py:354-356): sky - absorp / median(absorp) * m + m with m the median of the
            # central half of the (background-subtracted) spectrum; applied after the background like the reference
            sky = cube2d[index] if index is not None else cube2d
            if bkg_t is not None:
                sky = sky - bkg_t[None, :]
                bkg_t = None
            ab = torch.from_numpy(np.ascontiguousarray(absorp, dtype=np.float32)).to(dev)
            cut = int(nz * 0.25)
            sky = sky.contiguous()
            med = engine.row_median(sky, cut, nz - cut)                 # np.nanmedian(sky[cut:-cut]) per spaxel
            ab_med = float(np.nanmedian(np.asarray(absorp, dtype=np.float64)))      # (host input: one number)
            cube2d = (sky - (ab / ab_med)[None, :] * med[:, None] + med[:, None]).contiguous()
            index = None
        if auto_prior:
            v0, b0 = engine.estimate_prior(cfg, cube2d, theta_t, index=index, bkg=bkg_t)
            have_prior = True
        self.last_prior = (v0, b0)
        L_ = len(lines)
        K = cfg.row_floats
        # sharded single-pass fits store their rows straight into rank 0's buffers (dist.PeerRows); fits that compare
        # several passes per spaxel (component search, restarts) keep them local and gather the winners
        peer = None
        if sharded and not search and int(n_stoch) == 1:
            peer = self._peer_rows(n_fit, K, dev)
        if search:
            # per-group starts from the shared pair: the first group keeps it, every further group is offset
            n_loc = theta_t.numel()
            v_sh = v0.reshape(n_loc, 1) if v0 is not None else torch.zeros((n_loc, 1), dtype=torch.float32, device=dev)
            b_sh = b0.reshape(n_loc, 1) if b0 is not None else torch.zeros((n_loc, 1), dtype=torch.float32, device=dev)
            b_pg = b_sh.expand(n_loc, cfg.n_sigma_groups).contiguous()
            res = None
            for d in COMPONENT_OFFSETS:
                off = torch.zeros((1, cfg.n_vel_groups), dtype=torch.float32, device=dev)
                off[0, 1:] = d
                alt = engine.fit_batch(cfg, cube2d, theta_t, (v_sh + off).contiguous(), b_pg, index=index, bkg=bkg_t)
                res = alt if res is None else engine.keep_best(res, alt, 7 * L_)
            n_stoch = 1
        elif peer is not None:
            rows_w, status_w = peer.window(lo, hi)
            res = engine.fit_batch(cfg, cube2d, theta_t, v0, b0, index=index, bkg=bkg_t, rows=rows_w, status=status_w)
        else:
            res = engine.fit_batch(cfg, cube2d, theta_t, v0, b0, index=index, bkg=bkg_t)
        rows = res['rows']
        if have_prior:                                                  # (without a prior all draws would equal the start)
            res = self._restarts(cfg, cube2d, theta_t, v0, b0, res, n_stoch, n_fit, lo, hi, index=index, bkg=bkg_t)
        rows = res['rows']
        status = res['status']
        if sharded:
            rows, status = self._collect_rows(peer, rows, status, n_fit, lo, hi, dev)
        if sel is None:
            full = rows
        else:
            full = torch.zeros((nyo * nxo, K), dtype=torch.float32, device=dev)
            pos = torch.from_numpy(np.flatnonzero(sel.ravel())).to(dev)
            full[pos] = rows
        # one pass on the device to field-major [K, ny' nx'] (lucifit_pack_maps), one copy to the host: every 2-D map of a
        # line is then a contiguous slice (what save_fits writes), and the (ny', nx', L) arrays the API returns are views
        # of it.  Large maps go through pinned staging buffers kept on the object, and the writer also fetches the
        # byte-reversed copy the kernel made, from which the FITS files are written without a host-side byte swap.
        self._last_be = None
        n_map = int(full.shape[0])
        if n_map * K >= self.STAGE_MIN_VALUES:
            want_be = self._is_writer()
            packed = engine.pack_maps(full, want_be=want_be)
            maps_dev, be_dev = packed if want_be else (packed, None)
            stage = self._host_stage(K, n_map, want_be)
            stage['native'].copy_(maps_dev, non_blocking=True)
            if want_be:
                stage['be'].copy_(be_dev, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            host = np.empty((K, n_map), dtype=np.float32)           # the caller's arrays: never aliased with the staging
            src = stage['native'].numpy()
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
                list(pool.map(lambda i: np.copyto(host[i], src[i]), range(K)))
            if want_be:
                self._last_be = stage['be'].numpy()                  # valid until the next fit of this object
        else:
            host = engine.pack_maps(full).cpu().numpy()
        self.last_fit_summary = engine.status_summary(status)
        maps = {}
        for i, name in enumerate(engine.ROW_FIELDS):
            maps[name] = host[i * L:(i + 1) * L].reshape(L, nyo, nxo).transpose(1, 2, 0)
        for i, name in enumerate(engine.ROW_SCALARS):
            maps[name] = host[7 * L + i].reshape(nyo, nxo)
        return maps

    def _host_stage(self, K, n, want_be):
        """Pinned host staging for the maps of a large fit (``[K, n]`` float32 and, for the writer, ``[K, n]`` int32 holding
        big-endian floats), allocated once per shape and reused: a pageable copy of the 676 MB of a full-cube fit costs
        0.3 s, the pinned one 12 ms."""
        st = getattr(self, '_stage', None)
        if st is None or st['shape'] != (K, n):
            st = {'shape': (K, n), 'native': torch.empty((K, n), dtype=torch.float32, pin_memory=True), 'be': None}
            self._stage = st
        if want_be and st['be'] is None:
            st['be'] = torch.empty((K, n), dtype=torch.int32, pin_memory=True)
        return st

    @_nvtx('luci:save_maps')
    def _save_maps(self, lines, m, header, binning, shape, fit_function=None, suffix=''):
        """The FITS maps of a fit (LuciUtility.save_fits layout), from the big-endian staging copy when the fit left one."""
        if not self._is_writer():
            return
        be = getattr(self, '_last_be', None)
        if be is not None and be.shape[1] == shape[0] * shape[1]:
            from .fitsio import save_fits_be
            L = len(lines)
            index = {name: i * L for i, name in enumerate(engine.ROW_FIELDS)}
            index.update({name: 7 * L + i for i, name in enumerate(engine.ROW_SCALARS)})
            save_fits_be(self.output_dir, self.object_name, lines, be, lambda f, k: index[f] + (k or 0), shape, header,
                         binning, suffix=suffix, fit_function=fit_function)
            return
        save_fits(self.output_dir, self.object_name, lines, m['amplitudes'], m['fluxes'], m['flux_errors'],
                  m['velocities'], m['sigmas'], m['vels_errors'], m['sigmas_errors'], m['chi2'], m['continuum'],
                  m['continuum_error'], header, binning, suffix=suffix, fit_function=fit_function)

    def _mask_rows(self, mask_np):
        """int64 device row list (x-major, the order of the reference's spaxel loops) of a boolean ``(nx, ny)`` mask."""
        m = np.asarray(mask_np).astype(bool)
        nx, ny = self.cube_final.shape[:2]
        if m.shape != (nx, ny):
            raise ValueError('mask must be (nx, ny) = %s' % ((nx, ny),))
        return torch.from_numpy(np.flatnonzero(m.reshape(-1)).astype(np.int64)).to(self.cube_final.device)

    def _restarts(self, cfg, cube2d, theta_t, v0, b0, res, n_stoch, n_global, lo, hi, index=None, bkg=None, **kw):
        """Random restarts (``LuciFit.py:662-687``): the reference redraws every line's position from
        ``N(p, |1e7/lambda - p| / 50)`` and sigma from ``N(sigma, sigma / 10)`` with numpy's global generator and keeps the
        fit with the lowest loss; in velocity units that is ``v ~ N(v0, |v0| / 50)``, ``b ~ N(b0, b0 / 10)`` per tie
        group, drawn here with the same generator and compared on the chi2 each fit reports.  Every restart is one more
        batch call over the same spaxels.  The draws are made for the whole fitted list (``n_global``) and sliced to
        ``[lo, hi)``, so that a sharded call with equally seeded ranks makes the draws of the single-process call."""
        dev = theta_t.device
        for st in range(1, int(n_stoch)):
            v_np, b_np = v0.cpu().numpy().astype(np.float64), b0.cpu().numpy().astype(np.float64)
            wv, wb = v_np.size // max(hi - lo, 1), b_np.size // max(hi - lo, 1)
            zv = np.random.standard_normal((n_global, wv))[lo:hi].reshape(-1)
            zb = np.random.standard_normal((n_global, wb))[lo:hi].reshape(-1)
            v_r = torch.from_numpy((v_np + np.abs(v_np) / 50.0 * zv).astype(np.float32)).to(dev)
            b_r = torch.from_numpy((b_np + np.abs(b_np) / 10.0 * zb).astype(np.float32)).to(dev)
            alt = engine.fit_batch(cfg, cube2d, theta_t, v_r, b_r, index=index, bkg=bkg, **kw)
            res = engine.keep_best(res, alt, 7 * cfg.L)
        return res

    def _peer_rows(self, n_fit, K, dev):
        """The (cached) peer-mapped destination buffers of a sharded fit; ``None`` when the mapping is unavailable."""
        cache = getattr(self, '_peer_cache', None)
        if cache is None or cache[0] != (n_fit, K):
            if cache is not None:
                cache[1].close()
            self._peer_cache = cache = ((n_fit, K), ldist.PeerRows(n_fit, K, dev, dst=0))
        return cache[1] if cache[1].ok else None

    def _collect_rows(self, peer, rows, status, n_fit, lo, hi, dev):
        """Rows / status of all ranks on rank 0 (``[n_fit, K]``, ``[n_fit]``); the other ranks get their own run placed
        in zeros.  With ``peer`` the rows are already there once every rank's stream has drained."""
        rank = torch.distributed.get_rank()
        world = torch.distributed.get_world_size()
        if peer is not None:
            peer.complete()
            if rank == 0:
                return peer.rows, peer.status
            return (torch.zeros((n_fit, rows.shape[1]), dtype=torch.float32, device=dev),
                    torch.zeros((n_fit,), dtype=torch.int32, device=dev))       # (written remotely: nothing local to show)
        counts = [None] * world
        torch.distributed.all_gather_object(counts, hi - lo)
        all_rows = ldist.gather_rows(rows, dst=0, counts=counts)
        all_status = ldist.gather_rows(status.view(-1, 1), dst=0, counts=counts)
        if rank == 0:
            return all_rows, all_status.view(-1)
        full_r = torch.zeros((n_fit, rows.shape[1]), dtype=torch.float32, device=dev)
        full_s = torch.zeros((n_fit,), dtype=torch.int32, device=dev)
        full_r[lo:hi] = rows
        full_s[lo:hi] = status
        return full_r, full_s

    @staticmethod
    def _is_writer():
        """Rank 0 writes the FITS / NPY artefacts of a sharded call."""
        return (not ldist.is_distributed()) or torch.distributed.get_rank() == 0

    def _ensure_deep(self):
        if not self._is_writer():
            return
        if not os.path.exists(self.output_dir + '/' + self.object_name + '_deep.fits'):
            self.create_deep_image()

    def fit_cube(self, lines, fit_function, vel_rel, sigma_rel, x_min, x_max, y_min, y_max, bkg=None, bkgType=None,
                 binning=None, bayes_bool=False, bayes_method='emcee', absorp=None, uncertainty_bool=False,
                 n_threads=2, nii_cons=True, initial_values=[False], spec_min=None, spec_max=None, obj_redshift=0.0,
                 n_stoch=1, pca_coefficient_array=None, pca_vectors=None, pca_mean=None, prior_values=None):
        """``LuciBase.py:409-551``.  Returns ``(velocities, broadenings, flux, ampls)``, float32 ``(ny', nx', L)``.

        ``n_threads`` is accepted and ignored (the batch runs on the GPU).  ``prior_values=(vel, broad)`` are
        ``(ny', nx')`` maps in km/s replacing the reference's CNN prior."""
        self._reject_unsupported(bayes_bool, absorp, initial_values, n_stoch, bkgType)
        cube_to_slice = self.cube_final
        header = self.header
        x0_unbinned, y0_unbinned = x_min, y_min
        if binning is not None and binning != 1:
            self.bin_cube(self.cube_final, self.header, binning, x_min, x_max, y_min, y_max)
            x_max = int((x_max - x_min) / binning)
            y_max = int((y_max - y_min) / binning)
            x_min = 0
            y_min = 0
            cube_to_slice = self.cube_binned
            header = self.header_binned
        self._ensure_deep()
        use_bkg = bkg if bkgType == 'standard' else None                # LuciBase.py:325-330 (ignored otherwise)
        scale = (binning ** 2) if (binning and use_bkg is not None) else 1.0
        pca = None
        if bkgType == 'pca':
            if pca_coefficient_array is None or pca_vectors is None or pca_mean is None:
                raise Exception("bkgType='pca' needs pca_coefficient_array, pca_vectors and pca_mean")
            coef = np.asarray(pca_coefficient_array, dtype=np.float64)
            if binning is not None and binning != 1:                    # coefficients of a bin: nansum over it (:343-347)
                b_ = int(binning)
                c_ = coef[x0_unbinned:x0_unbinned + x_max * b_, y0_unbinned:y0_unbinned + y_max * b_]
                coef = np.nansum(np.nansum(c_.reshape(x_max, b_, y_max, b_, -1), axis=3), axis=1)
            else:
                coef = coef[x_min:x_max, y_min:y_max]
            pca = (coef, np.asarray(pca_vectors), np.asarray(pca_mean))
        elif bkgType not in (None, 'standard'):
            raise Exception('Please set bkgType to either standard or pca')
        m = self._fit_maps(cube_to_slice, lines, fit_function, vel_rel, sigma_rel, x_min, x_max, y_min, y_max, None,
                           use_bkg, scale, uncertainty_bool, nii_cons, spec_min, spec_max, obj_redshift, prior_values,
                           n_stoch=n_stoch, pca=pca, initial_values=initial_values, absorp=absorp)
        self._save_maps(lines, m, self._cutout_header(header, x_min, y_min, x_max, y_max), binning,
                        m['chi2'].shape, fit_function=fit_function)
        self.last_fit_maps = m
        return m['velocities'], m['sigmas'], m['fluxes'], m['amplitudes']

    def fit_entire_cube(self, lines, fit_function, vel_rel, sigma_rel, bkg=None, binning=None, bayes_bool=False,
                        output_name=None, uncertainty_bool=False, n_threads=1, **kwargs):
        """``LuciBase.py:231-255``; unlike the reference the keyword arguments are honoured and the maps returned."""
        return self.fit_cube(lines, fit_function, vel_rel, sigma_rel, 0, self.cube_final.shape[0], 0,
                             self.cube_final.shape[1], bkg=bkg, binning=binning, bayes_bool=bayes_bool,
                             uncertainty_bool=uncertainty_bool, n_threads=n_threads, **kwargs)

    def close(self):
        """``LuciBase.py:1302-1309``: drop the cube (here: its device memory) and the header."""
        self.cube_final = None
        self.header = None
        self.cube_binned = None
        if torch.cuda.is_available():
            torch.cuda.empty_cache()

    def _load_mask(self, region, pixel_list):
        if pixel_list is not False:
            mask = np.ones((self.cube_final.shape[0], self.cube_final.shape[1]), dtype=bool)   # LuciBase.py:637-639
            for pair in region:
                mask[pair] = True
            return mask
        if isinstance(region, str):
            if '.reg' in region:                        # LuciConvenience.reg_to_mask :49-57 (pyregion restated, regions.py)
                from .regions import reg_to_mask
                return reg_to_mask(region, self.header, self.cube_final.shape)
            if '.npy' in region:
                return np.load(region)
            raise ValueError('Mask was incorrectly passed. Please use either a .reg file or a .npy file or a numpy ndarray')
        if region is not None:
            return np.asarray(region)
        raise ValueError('Mask was incorrectly passed. Please use either a .reg file or a .npy file or a numpy ndarray')

    def fit_region(self, lines, fit_function, vel_rel, sigma_rel, region, bkg=None, binning=None, bayes_bool=False,
                   bayes_method='emcee', output_name=None, uncertainty_bool=False, n_threads=1, nii_cons=True,
                   spec_min=None, spec_max=None, obj_redshift=0.0, initial_values=[False], n_stoch=1,
                   pixel_list=False, prior_values=None):
        """``LuciBase.py:553-722``.  Returns ``(velocities, broadenings, flux, chi2, mask)``.  Only the spaxels inside
        the mask are sent to the GPU (compacted index list); the rest of every map stays 0 like the reference."""
        self._reject_unsupported(bayes_bool, None, initial_values, n_stoch, None)
        x_min, x_max = 0, self.cube_final.shape[0]
        y_min, y_max = 0, self.cube_final.shape[1]
        cube_to_slice = self.cube_final
        header = self.header
        if binning is not None and binning > 1:
            self.bin_cube(self.cube_final, self.header, binning, x_min, x_max, y_min, y_max)
            x_max = int((x_max - x_min) / binning)
            y_max = int((y_max - y_min) / binning)
            cube_to_slice = self.cube_binned
            header = self.header_binned
        mask = self._load_mask(region, pixel_list)
        if binning is not None and binning > 1:
            from .host_ops import bin_mask
            mask = bin_mask(mask, binning, x_min, self.cube_final.shape[0], y_min, self.cube_final.shape[1])
        self._ensure_deep()
        # fit_region never forwards bkgType, so the reference ignores bkg here (LuciBase.py:691-705, 325)
        m = self._fit_maps(cube_to_slice, lines, fit_function, vel_rel, sigma_rel, x_min, x_max, y_min, y_max, mask,
                           None, 1.0, uncertainty_bool, nii_cons, spec_min, spec_max, obj_redshift, prior_values,
                           n_stoch=n_stoch, initial_values=initial_values)
        self._save_maps(lines, m, self._cutout_header(header, x_min, y_min, x_max, y_max), binning, m['chi2'].shape)
        self.last_fit_maps = m
        return m['velocities'], m['sigmas'], m['fluxes'], m['chi2'], mask

    def fit_spectrum_region(self, lines, fit_function, vel_rel, sigma_rel, region, initial_values=[False], bkg=None,
                            bayes_bool=False, bayes_method='emcee', uncertainty_bool=False, mean=False,
                            nii_cons=True, spec_min=None, spec_max=None, obj_redshift=0.0, n_stoch=1,
                            prior_values=None):
        """``LuciBase.py:949-1046``: fit the spectrum integrated over a mask.  Returns ``(axis, sky, fit_dict)``.
        ``prior_values=(vel, broad)`` are scalars in km/s; ``initial_values=[vel, broad]`` holds them fixed (the freeze
        branch, ``LuciFit.py:688-709``); ``n_stoch`` restarts like ``LuciFit.py:662-687``.  NaN channels of the
        integrated spectrum are left out of the fit like the reference's ``good_sky_inds`` (``:1027-1030``)."""
        self._reject_unsupported(bayes_bool, None, initial_values, n_stoch, None)
        rows = self._mask_rows(self._load_mask(region, False))
        spec_ct = int(rows.numel())
        nx_, ny_, nz_ = self.cube_final.shape
        integrated = engine.segment_sum(self.cube_final.view(nx_ * ny_, nz_), rows)[0]      # `+=` of :1019: NaN poisons a channel
        if mean:
            integrated = integrated / spec_ct
        if bkg is not None:
            integrated = integrated - torch.from_numpy(np.asarray(bkg, dtype=np.float64)).to(integrated.device) * spec_ct
        # the reference passes the angle of the LAST pixel of its double loop (LuciBase.py:1035)
        theta = float(self.interferometer_theta[self.cube_final.shape[0] - 1, self.cube_final.shape[1] - 1])
        fit_dict = self._fit_one_spectrum(integrated, theta, lines, fit_function, vel_rel, sigma_rel, uncertainty_bool,
                                          nii_cons, spec_min, spec_max, obj_redshift, prior_values, n_stoch=n_stoch,
                                          initial_values=initial_values)
        sky = integrated.cpu().numpy()
        return np.asarray(self.spectrum_axis), sky, fit_dict

    def _fit_one_spectrum(self, spectrum, theta, lines, fit_function, vel_rel, sigma_rel, uncertainty_bool, nii_cons,
                          spec_min, spec_max, obj_redshift, prior_values, n_stoch=1, initial_values=(False,)):
        """One spectrum (device tensor ``[nz]``) through ``lucifit_fit_batch``; returns the dictionary of
        ``Fit.fit`` (``LuciFit.py:824-834``).  Shared by fit_spectrum_region / fit_pixel."""
        sky32 = spectrum.to(torch.float32).view(1, -1).contiguous()
        freeze = len(initial_values) >= 2 and initial_values[0] is not False
        if freeze:
            if uncertainty_bool:
                raise NotImplementedError('uncertainty_bool with initial_values is broken in the reference '
                                          '(LuciFit.py:713-722 on an L+1 vector) and not offered')
            prior_values = (float(np.asarray(initial_values[0]).reshape(-1)[0]), float(np.asarray(initial_values[1]).reshape(-1)[0]))
            n_stoch = 1
        have_prior = prior_values is not None
        auto_prior = (not have_prior) and bool(self.ML_bool)
        cfg = self._make_config(lines, fit_function, vel_rel, sigma_rel, uncertainty_bool, nii_cons, spec_min, spec_max,
                                obj_redshift, ml_scale=have_prior or auto_prior, freeze=freeze)
        dev = sky32.device
        th = torch.tensor([theta], dtype=torch.float32, device=dev)
        v0 = torch.tensor([float(prior_values[0])], dtype=torch.float32, device=dev) if have_prior else None
        b0 = torch.tensor([float(prior_values[1])], dtype=torch.float32, device=dev) if have_prior else None
        if auto_prior:                                          # the reference's CNN prior, measured on the device
            v0, b0 = engine.estimate_prior(cfg, sky32, th)
            prior_values = (float(v0[0].item()), float(b0[0].item()))
            have_prior = True
        res = engine.fit_batch(cfg, sky32, th, v0, b0, want_sol=True, want_unc=True)
        if have_prior and int(n_stoch) > 1:
            res = self._restarts(cfg, sky32, th, v0, b0, res, n_stoch, 1, 0, 1, want_sol=True, want_unc=True)
        L = len(lines)
        row = res['rows'][0].cpu().numpy()
        q = engine.unpack_rows(row[None, :], L)
        fit_dict = {'fit_sol': res['fit_sol'][0].cpu().numpy().astype(np.float64),
                    'fit_uncertainties': res['fit_unc'][0].cpu().numpy().astype(np.float64),
                    'amplitudes': list(q['amplitudes'][0]), 'fluxes': list(q['fluxes'][0]),
                    'flux_errors': list(q['flux_errors'][0]), 'chi2': float(q['chi2'][0]),
                    'velocities': list(q['velocities'][0]), 'sigmas': list(q['sigmas'][0]),
                    'vels_errors': list(q['vels_errors'][0]), 'sigmas_errors': list(q['sigmas_errors'][0]),
                    'axis_step': float(q['axis_step'][0]), 'corr': float(q['corr'][0]),
                    'continuum': float(q['continuum'][0]), 'continuum_error': float(q['continuum_error'][0]),
                    'flat_samples': None, 'vel_ml': float(prior_values[0]) if have_prior else 0.0, 'vel_ml_sigma': 0,
                    'broad_ml': abs(float(prior_values[1])) if have_prior else 0.0, 'broad_ml_sigma': 0,
                    'fit_axis': np.asarray(self.spectrum_axis) * (1 + obj_redshift),
                    'status': int(res['status'][0].item())}
        fit_dict['scale'] = None
        from .host_ops import plot_vector
        fit_dict['fit_vector'] = plot_vector(cfg, fit_dict['fit_sol'], theta, self.cube_final.device)
        return fit_dict

    def fit_pixel(self, lines, fit_function, vel_rel, sigma_rel, pixel_x, pixel_y, binning=None, bkg=None, absorp=None,
                  bayes_bool=False, bayes_method='emcee', uncertainty_bool=False, nii_cons=True, spec_min=None,
                  spec_max=None, obj_redshift=0.0, n_stoch=1, bkgType='standard', pca_coefficient_array=None,
                  pca_vectors=None, pca_mean=None, prior_values=None):
        """``LuciBase.py:724-838``: fit one pixel, or with ``binning`` the nansum of the ``2 binning`` square
        ``[pixel - binning, pixel + binning)`` around it.  Returns ``(axis, sky, fit_dict)``.

        Like the reference the background is scaled by ``binning ** 2`` (``:769``), which makes ``bkg`` usable only
        together with ``binning`` there (``None ** 2`` raises); here ``binning=None`` subtracts the plain background.
        The reference subtracts in place from a view of ``cube_final`` (``:767-769``), altering the cube; that side
        effect is not reproduced.  ``prior_values=(vel, broad)`` are scalars in km/s."""
        self._reject_unsupported(bayes_bool, absorp, [False], n_stoch, bkgType)
        if absorp is not None:
            raise NotImplementedError('absorp is implemented for fit_cube only')
        b = int(binning) if (binning is not None and binning != 1) else 0
        if b:
            nx_, ny_, nz_ = self.cube_final.shape
            dev_ = self.cube_final.device
            xs = torch.arange(max(pixel_x - b, 0), min(pixel_x + b, nx_), dtype=torch.int64, device=dev_)
            ys = torch.arange(max(pixel_y - b, 0), min(pixel_y + b, ny_), dtype=torch.int64, device=dev_)
            sky = engine.segment_sum(self.cube_final.view(nx_ * ny_, nz_), (xs[:, None] * ny_ + ys[None, :]).reshape(-1),
                                     nansum=True)[0]
        else:
            sky = self.cube_final[pixel_x, pixel_y, :].to(torch.float64)
        if bkgType == 'standard':
            if bkg is not None:
                scale = float(binning) ** 2 if binning is not None else 1.0
                sky = sky - torch.from_numpy(np.asarray(bkg, dtype=np.float64)).to(sky.device) * scale
        elif bkgType == 'pca':
            if pca_coefficient_array is None or pca_vectors is None or pca_mean is None:
                raise Exception("bkgType='pca' needs pca_coefficient_array, pca_vectors and pca_mean")
            band = {'SN3': (675.0, 670.0), 'SN2': (505.0, 480.0)}.get(self.hdr_dict['FILTER'])
            if band is None:
                raise Exception('the PCA background is implemented for SN3 and SN2 (LuciBase.py:774-783)')
            nm_axis = 1e7 / np.asarray(self.spectrum_axis, dtype=np.float64)
            lo_i, hi_i = int(np.argmin(np.abs(nm_axis - band[0]))), int(np.argmin(np.abs(nm_axis - band[1])))
            coef = np.asarray(pca_coefficient_array, dtype=np.float64)
            if b:                                                       # coefficients of the box: nansum (:784-788)
                c = np.nansum(coef[max(pixel_x - b, 0):pixel_x + b, max(pixel_y - b, 0):pixel_y + b, :], axis=(0, 1))
            else:
                c = coef[pixel_x, pixel_y]                              # (the reference names an undefined x_pix here, :790)
            bkg_pix = np.asarray(pca_mean, dtype=np.float64) + c @ np.asarray(pca_vectors, dtype=np.float64)
            ref_band = sky[lo_i:hi_i]
            scale_spec = ref_band[~torch.isnan(ref_band)].max()
            sky = sky - scale_spec * torch.from_numpy(bkg_pix).to(sky.device)
        else:
            raise Exception('Please set bkgType to either standard or pca')
        # NaN channels are left out of the fit like the reference's good_sky_inds (LuciBase.py:801-803)
        theta = float(self.interferometer_theta[pixel_x, pixel_y])
        fit_dict = self._fit_one_spectrum(sky, theta, lines, fit_function, vel_rel, sigma_rel, uncertainty_bool, nii_cons,
                                          spec_min, spec_max, obj_redshift, prior_values, n_stoch=n_stoch)
        return np.asarray(self.spectrum_axis), sky.cpu().numpy(), fit_dict

    def extract_spectrum(self, x_min, x_max, y_min, y_max, bkg=None, binning=None, mean=False):
        """``LuciBase.py:844-893``: spectrum summed over a rectangle (NaN channels skipped), ``mean`` divides by the
        reference's ``spec_ct`` -- which it increments only ONCE (``:887-889``), so ``mean=True`` returns the plain sum
        there; with ``strict_reference=False`` the mean over the summed spectra is returned instead.  The background
        is subtracted from every summed spectrum (times ``binning ** 2`` when binned, ``:882-886``)."""
        cube = self.cube_final
        if binning is not None and binning != 1:
            self.bin_cube(self.cube_final, self.header, binning, x_min, x_max, y_min, y_max)
            x_max = int((x_max - x_min) / binning)
            y_max = int((y_max - y_min) / binning)
            x_min = y_min = 0
            cube = self.cube_binned
        nxc_, nyc_, nz_ = cube.shape
        xs = torch.arange(max(x_min, 0), min(x_max, nxc_), dtype=torch.int64, device=cube.device)
        ys = torch.arange(max(y_min, 0), min(y_max, nyc_), dtype=torch.int64, device=cube.device)
        rows = (xs[:, None] * nyc_ + ys[None, :]).reshape(-1)
        n_spec = int(rows.numel())
        integrated = engine.segment_sum(cube.view(nxc_ * nyc_, nz_), rows, nansum=True)[0]
        if bkg is not None:
            scale = float(binning) ** 2 if binning else 1.0
            integrated = integrated - torch.from_numpy(np.asarray(bkg, dtype=np.float64)).to(integrated.device) * (scale * n_spec)
        if mean and not self.strict_reference:
            integrated = integrated / max(n_spec, 1)
        return np.asarray(self.spectrum_axis), integrated.cpu().numpy()

    def extract_spectrum_region(self, region, mean=False):
        """``LuciBase.py:895-947``: spectrum summed (or averaged) over a boolean mask / ``.npy`` mask."""
        rows = self._mask_rows(self._load_mask(region, False))
        spec_ct = int(rows.numel())
        nx_, ny_, nz_ = self.cube_final.shape
        integrated = engine.segment_sum(self.cube_final.view(nx_ * ny_, nz_), rows)[0]
        if mean:
            integrated = integrated / spec_ct
        return np.asarray(self.spectrum_axis), integrated.cpu().numpy()

    def skyline_calibration(self, Luci_path, n_grid, bin_size=30):
        """``LuciBase.py:1205-1284``: velocity offset of the OH sky line(s) of ``<Luci_path>/Data/sky_lines.dat``
        (6498.729 Angstrom) on an ``n_grid x n_grid`` set of regions, fitted with a sinc started at 80 km/s
        (``Fit.fit(sky_line=True)``, ``LuciFit.py:837-928``).  Returns ``(velocity, fit_vector, sky, vel_grid,
        vel_uncertainty_grid, spectrum_axis)`` like the reference (the first three belong to the last region) and
        writes ``velocity_correction.fits``.

        The reference fits the regions one after the other; here their spectra are one gathered reduction and ONE
        ``lucifit_fit_batch`` call.  Region spectra: the reference adds ``cube_final[x_c + i, y_c + i]`` ``bin_size``
        times for ``i < bin_size`` (its inner index ``j`` is unused, ``:1254-1256``), i.e. ``bin_size`` times the
        diagonal of the box; ``strict_reference=False`` sums the full ``bin_size x bin_size`` box instead.  Every line
        of the file gets its own free position like in the reference (no velocity tie, ``LuciFit.py:866``); the
        returned velocity is that of the first line."""
        import csv
        with open(str(Luci_path).rstrip('/') + '/Data/sky_lines.dat') as fh:
            rows = [r for r in csv.reader(fh) if r and not r[0].startswith('#')]
        names = [c.strip() for c in rows[0]]
        waves_nm = [float(r[names.index('Wavelength')]) / 10.0 for r in rows[1:]]      # Angstrom -> nm (:1229)
        if not 1 <= len(waves_nm) <= 8:
            raise NotImplementedError('skyline_calibration: 1 to 8 sky lines per fit are supported')
        sky_lines = {'OH_%i' % k: w for k, w in enumerate(waves_nm)}
        nx, ny, nz = self.cube_final.shape
        x_min, y_min = 200, 200                                                        # :1236-1241
        x_max, y_max = nx - x_min, ny - y_min
        x_step, y_step = int((x_max - x_min) / n_grid), int((y_max - y_min) / n_grid)
        dev = self.cube_final.device
        centres = [(x_min + int(0.5 * x_step * (xg + 1)), y_min + int(0.5 * y_step * (yg + 1)))
                   for xg in range(n_grid) for yg in range(n_grid)]
        flat = self.cube_final.view(nx * ny, nz)
        if self.strict_reference:
            offs = torch.arange(bin_size, device=dev) * (ny + 1)                       # (x_c + i, y_c + i)
            scale = float(bin_size)
        else:
            ii, jj = torch.meshgrid(torch.arange(bin_size, device=dev), torch.arange(bin_size, device=dev), indexing='ij')
            offs = (ii * ny + jj).reshape(-1)
            scale = 1.0
        base = torch.tensor([xc * ny + yc for xc, yc in centres], dtype=torch.int64, device=dev)
        seg = torch.arange(len(centres), dtype=torch.int32, device=dev).repeat_interleave(offs.numel())
        spectra = engine.segment_sum(flat, (base[:, None] + offs[None, :]).reshape(-1), seg, len(centres)) * scale
        theta = np.array([self.interferometer_theta[xc, yc] for xc, yc in centres], dtype=np.float32)
        L = len(waves_nm)
        cfg = FitConfig('sinc', list(sky_lines), list(range(1, L + 1)), [1] * L, self.spectrum_axis,
                        filter=self.hdr_dict['FILTER'], delta_x=self.hdr_dict['STEP'], n_steps=self.step_nb,
                        zpd_index=self.zpd_index, trans_filter=self.transmission_interpolated, uncertainty_bool=True,
                        nii_cons=False, ml_scale=False, prior_groups=L > 1, extra_lines=sky_lines)
        n = len(centres)
        # The reference starts every position at 80 km/s (:859) and lets SLSQP walk; the sinc objective has a local
        # minimum per side lobe, so the start is first moved to the maximum of the matched filter within +-200 km/s of
        # that guess (lucifit_estimate_prior) -- the same optimum, reached from inside its basin.
        spectra32 = spectra.to(torch.float32).contiguous()
        theta_t = torch.from_numpy(theta).to(dev)
        v_start, _ = engine.estimate_prior(cfg, spectra32, theta_t, v_range=(SKYLINE_V0 - 200.0, SKYLINE_V0 + 200.0))
        # per-group priors when L > 1: one velocity per line (each line is its own velocity group), one broadening per
        # sigma group (a single one: sigma_rel = [1] * L)
        v0 = v_start.repeat_interleave(cfg.n_vel_groups).contiguous() if L > 1 else v_start
        b0 = torch.zeros((n * (cfg.n_sigma_groups if L > 1 else 1),), dtype=torch.float32, device=dev)
        res = engine.fit_batch(cfg, spectra32, theta_t, v0, b0, want_sol=True)
        q = engine.unpack_rows(res['rows'].cpu().numpy(), L)
        vel_grid = q['velocities'][:, 0].astype(np.float64).reshape(n_grid, n_grid)
        # the sky-line branch inverts the Hessian itself (LuciFit.py:873-876), the regular fit half of it (:718): its
        # 1-sigma errors are smaller by sqrt(2)
        vel_uncertainty_grid = (q['vels_errors'][:, 0].astype(np.float64) / np.sqrt(2.0)).reshape(n_grid, n_grid)
        write_fits(self.output_dir + '/velocity_correction.fits', vel_grid, wcs_to_header(self.header))
        from .host_ops import plot_vector
        fit_vector = plot_vector(cfg, res['fit_sol'][-1].cpu().numpy().astype(np.float64), float(theta[-1]), dev)
        self.last_fit_summary = engine.status_summary(res['status'])
        return (float(vel_grid[-1, -1]), fit_vector, spectra[-1].cpu().numpy(), vel_grid, vel_uncertainty_grid,
                np.asarray(self.spectrum_axis))

    def fit_wvt(self, lines, fit_function, vel_rel, sigma_rel, bkg=None, bayes_bool=False, uncertainty_bool=False,
                n_threads=1, initial_values=[False], n_stoch=1, stn_target=10, bin_map=None, prior_values=None,
                nii_cons=True, mean=False):
        """``LuciBase.py:1383-1467``: fit the integrated spectrum of every bin of a tessellation and paint each bin's
        result onto its pixels.  Returns ``(velocities, broadenings, flux, chi2, header)`` like the reference.

        The reference loops over ``<output_dir>/Numpy_Voronoi_Bins/bool_bin_map_<i>.npy`` and calls
        ``fit_spectrum_region`` once per bin (9.5 h for 16 789 bins, docs/source/example_wvt.rst:116).  Here the bins
        are one segmented reduction of the cube followed by ONE ``lucifit_fit_batch`` call.  ``bin_map`` (int
        ``(nx, ny)``, -1 = unbinned pixel) may be passed directly instead of the ``.npy`` files; the tessellation
        itself (``create_wvt``, LuciWVT.py) is out of scope.  ``prior_values=(vel_map, broad_map)`` ``(ny, nx)`` are
        reduced to one prior per bin by the median over its pixels; ``initial_values=[vel_map, broad_map]`` (indexed
        ``[x, y]`` like the reference does) freezes every bin at the values of its last pixel; ``n_stoch`` restarts are
        further batch calls over all bins.  Every bin is fitted with the interferometer angle
        the reference's ``fit_spectrum_region`` uses (LuciBase.py:1035: the last pixel of the cube).

        Two defects of the reference loop are not reproduced: its maps are ``(ny, nx)`` but indexed ``[x, y]``
        (IndexError off the square part of the detector, LuciBase.py:1449-1459) and the continuum map is overwritten
        with the continuum error (:1456-1457)."""
        self._reject_unsupported(bayes_bool, None, [False], n_stoch, None)
        freeze = len(initial_values) >= 2 and initial_values[0] is not False
        if freeze and uncertainty_bool:
            raise NotImplementedError('uncertainty_bool with initial_values is broken in the reference '
                                      '(LuciFit.py:713-722 on an L+1 vector) and not offered')
        nx, ny, nz = self.cube_final.shape
        dev = self.cube_final.device
        if bin_map is None:
            folder = os.path.join(self.output_dir, 'Numpy_Voronoi_Bins')
            bin_map = np.full((nx, ny), -1, dtype=np.int64)
            for b in range(len(os.listdir(folder))):
                bin_map[np.load(os.path.join(folder, 'bool_bin_map_%i.npy' % b)).astype(bool)] = b
        bin_map = np.asarray(bin_map, dtype=np.int64)
        if bin_map.shape != (nx, ny):
            raise ValueError('bin_map must be (nx, ny) = %s' % ((nx, ny),))
        n_bins = int(bin_map.max()) + 1
        if n_bins < 1:
            raise ValueError('bin_map holds no bins')
        flat = torch.from_numpy(bin_map.reshape(-1)).to(dev)
        inside = flat >= 0
        # segmented reduction: integrated spectrum of every bin in one pass (lucifit_segment_sum, float64 sums); the pixel
        # list is ordered by bin so that a CTA's running sums are flushed once per bin
        order = torch.sort(flat, stable=True)[1]
        order = order[int((~inside).sum().item()):]                 # (unbinned pixels carry -1 and sort first)
        sums = engine.segment_sum(self.cube_final.view(nx * ny, nz), order.contiguous(), flat[order].to(torch.int32).contiguous(),
                                  n_bins, nansum=True)
        counts = torch.bincount(flat[inside], minlength=n_bins).to(torch.float64)
        if mean:
            sums = sums / counts.clamp(min=1.0)[:, None]
        if bkg is not None:
            sums = sums - torch.from_numpy(np.asarray(bkg, dtype=np.float64)).to(dev)[None, :] * counts[:, None]
        theta = float(self.interferometer_theta[nx - 1, ny - 1])
        have_prior = prior_values is not None or freeze
        auto_prior = (not have_prior) and bool(self.ML_bool)
        cfg = self._make_config(lines, fit_function, vel_rel, sigma_rel, uncertainty_bool, nii_cons, None, None, 0.0,
                                ml_scale=have_prior or auto_prior, freeze=freeze)
        th = torch.full((n_bins,), theta, dtype=torch.float32, device=dev)
        v0 = b0 = None
        if freeze:
            # initial_values = [velocity map, broadening map] indexed [x, y]: every bin is held at the values of its LAST
            # pixel in np.where order (the reference's inner loop leaves them there, LuciBase.py:1437-1442)
            vi, bi = np.asarray(initial_values[0], dtype=np.float32), np.asarray(initial_values[1], dtype=np.float32)
            fm = bin_map.reshape(-1)
            last_pix = np.full(n_bins, -1, dtype=np.int64)
            last_pix[fm[fm >= 0]] = np.flatnonzero(fm >= 0)              # later pixels overwrite earlier ones
            has = last_pix >= 0
            vb = np.zeros((2, n_bins), np.float32)
            vb[0, has] = vi.reshape(-1)[last_pix[has]]
            vb[1, has] = bi.reshape(-1)[last_pix[has]]
            v0, b0 = torch.from_numpy(vb[0].copy()).to(dev), torch.from_numpy(vb[1].copy()).to(dev)
            n_stoch = 1
        elif have_prior:
            pv = np.asarray(prior_values[0], dtype=np.float64).T.reshape(-1)     # (ny, nx) -> x-major like bin_map
            pb = np.asarray(prior_values[1], dtype=np.float64).T.reshape(-1)
            order = np.argsort(bin_map.reshape(-1), kind='stable')
            keys = bin_map.reshape(-1)[order]
            first = np.searchsorted(keys, np.arange(n_bins), side='left')
            last = np.searchsorted(keys, np.arange(n_bins), side='right')
            med = lambda a: np.array([np.median(a[order[f:l]]) if l > f else 0.0 for f, l in zip(first, last)], np.float32)
            v0, b0 = torch.from_numpy(med(pv)).to(dev), torch.from_numpy(med(pb)).to(dev)
        sums32 = sums.to(torch.float32).contiguous()
        if auto_prior:
            v0, b0 = engine.estimate_prior(cfg, sums32, th)
        res = engine.fit_batch(cfg, sums32, th, v0, b0)
        if v0 is not None and int(n_stoch) > 1:
            res = self._restarts(cfg, sums32, th, v0, b0, res, n_stoch, n_bins, 0, n_bins)
        self.last_fit_summary = engine.status_summary(res['status'])
        L, K = len(lines), cfg.row_floats
        # paint: pixel (x, y) of bin b gets row b; unbinned pixels stay 0
        full = torch.zeros((nx * ny, K), dtype=torch.float32, device=dev)
        full[inside] = res['rows'][flat[inside]]
        full = full.view(nx, ny, K).permute(1, 0, 2).contiguous().cpu().numpy()      # (ny, nx, K) like every other map
        m = {}
        for i, name in enumerate(engine.ROW_FIELDS):
            m[name] = np.ascontiguousarray(full[:, :, i * L:(i + 1) * L])
        for i, name in enumerate(engine.ROW_SCALARS):
            m[name] = np.ascontiguousarray(full[:, :, 7 * L + i])
        m['bin_rows'] = res['rows'].cpu().numpy()
        self._ensure_deep()
        header = self._cutout_header(self.header, 0, 0)
        save_fits(self.output_dir, self.object_name, lines, m['amplitudes'], m['fluxes'], m['flux_errors'],
                  m['velocities'], m['sigmas'], m['vels_errors'], m['sigmas_errors'], m['chi2'], m['continuum'],
                  m['continuum_error'], header, binning=1, suffix='_wvt_%i' % stn_target)
        self.last_fit_maps = m
        return m['velocities'], m['sigmas'], m['fluxes'], m['chi2'], header
