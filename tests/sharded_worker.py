"""Worker of tests/test_gpu_sharded_api.py: one of N processes of a sharded ``Luci`` call (torchrun-style environment).

All processes may share ONE GPU (the driver's single-GPU box): the control plane is then gloo, the data plane -- rows
written straight into rank 0's buffers through ``lucifit_peer_open`` -- is the same as on N GPUs under NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SN3 = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731']


def run(out_dir, sharded):
    from luci_b200.luci import Luci
    from luci_b200.sim import SyntheticCube
    n_dev = torch.cuda.device_count()
    rank = int(os.environ.get('RANK', '0'))
    dev = torch.device('cuda', rank % n_dev if sharded else 0)
    torch.cuda.set_device(dev)
    sc = SyntheticCube(14, 11, model='sincgauss', resolution=5000, seed=5, snr_range=(20.0, 100.0))
    cube = sc.synthesize(dev, seed=77)
    luci = Luci.from_arrays(cube, sc.hdr_dict, os.path.join(out_dir, 'r%d' % rank), 'OBJ', spectrum_axis=sc.axis32, device=dev)
    v0, b0 = sc.priors()
    res = {}
    vel, broad, flux, ampl = luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 1, 13, 0, 11,
                                           prior_values=(v0[0:11, 1:13], b0[0:11, 1:13]), uncertainty_bool=True)
    res.update(cube_vel=vel, cube_flux=flux, cube_chi2=luci.last_fit_maps['chi2'], cube_verr=luci.last_fit_maps['vels_errors'])
    res['cube_peer'] = np.array(getattr(luci, '_peer_cache', (None, None))[1] is not None
                                and luci._peer_cache[1].ok)
    res['cube_n_converged'] = np.array(luci.last_fit_summary['n_converged'])
    mask = np.zeros((14, 11), bool)
    mask[2:9, 3:10] = True
    mask[5, 5] = False
    vel, broad, flux, chi2, _ = luci.fit_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, prior_values=(v0, b0))
    res.update(region_vel=vel, region_flux=flux, region_chi2=chi2)
    # restarts keep their rows local and gather the winners (torch.distributed point-to-point)
    np.random.seed(11)
    vel, broad, flux, ampl = luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 0, 14, 0, 11, prior_values=(v0, b0), n_stoch=2)
    res.update(stoch_vel=vel, stoch_chi2=luci.last_fit_maps['chi2'])
    return res


def main():
    out_dir, sharded = sys.argv[1], sys.argv[2] == '1'
    if sharded:
        backend = 'nccl' if torch.cuda.device_count() >= int(os.environ['WORLD_SIZE']) else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        dist.init_process_group(backend=backend)
    res = run(out_dir, sharded)
    rank = dist.get_rank() if sharded else 0
    np.savez(os.path.join(out_dir, 'maps_%s_rank%d.npz' % ('sharded' if sharded else 'single', rank)), **res)
    if sharded:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
