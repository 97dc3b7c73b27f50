"""The drop-in call under ``torchrun``: ``Luci.fit_cube`` / ``fit_region`` sharded over processes must return, on rank
0, exactly the maps of the single-process call (the reference's joblib fan-out, LuciBase.py:514-530, changes nothing
about the numbers either).  Runs with 2 processes on however many GPUs the box has (1 is enough: gloo control plane,
peer-mapped rows)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_sharded_maps_equal_single_process_maps(tmp_path):
    worker = os.path.join(ROOT, 'tests', 'sharded_worker.py')
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, worker, str(tmp_path), '0'], check=True, env=env, timeout=600)
    # (rows that go to a range mapped by lucifit_peer_open leave the output kernel through its staged, coalesced store path:
    # this comparison is also the bit-for-bit check of that path against the direct stores of the single-process run)
    run = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), worker, str(tmp_path), '1'],
                         check=True, env=dict(env, LUCIFIT_DEBUG='1'), timeout=900, stderr=subprocess.PIPE, text=True)
    assert 'stage_rows=1' in run.stderr and 'stage_rows=0' in run.stderr, run.stderr[-2000:]   # rank 1 staged, rank 0 direct
    one = np.load(os.path.join(tmp_path, 'maps_single_rank0.npz'))
    two = np.load(os.path.join(tmp_path, 'maps_sharded_rank0.npz'))
    assert bool(two['cube_peer']), 'rows were not written through peer memory'
    for key in one.files:
        if key == 'cube_peer':
            continue
        np.testing.assert_array_equal(one[key], two[key], err_msg=key)
    # rank 1 wrote no files and holds only its own share
    assert not os.path.exists(os.path.join(tmp_path, 'r1', 'Luci_outputs', 'OBJ_sincgauss_Chi2.fits'))
    assert os.path.exists(os.path.join(tmp_path, 'r0', 'Luci_outputs', 'OBJ_sincgauss_Chi2.fits'))
