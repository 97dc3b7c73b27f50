"""Sharding of the spaxel list over GPUs of one node and the collection of the output maps on one rank.

The fit path has no exchange step (every spaxel is independent; the reference's unit of parallelism is a y-row
handed to a joblib worker, ``LuciBase.py:514-517``).  Each rank owns a contiguous run of the fitted spaxels, fits it
with no collective, and only the ``(7L+5)`` float32 output columns have to reach the rank that assembles the maps:

* :class:`PeerRows` (the GPU path): the destination rank's map buffer is mapped into every other process
  (``lucifit_peer_export`` / ``lucifit_peer_open``) and each rank's output kernel stores its rows straight into it
  over NVLink -- there is no gather step, the transfer rides inside the fit;
* :func:`gather_rows`: point-to-point gather through ``torch.distributed`` (NCCL on the GPU, gloo in the CPU tests),
  used where peer mapping is unavailable and for fits that post-process their rows locally (restarts).
"""
import ctypes as C
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init_process_group(backend=None):
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29577')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def bind_to_gpu_cpus(local_rank):
    """Pin this process to the CPU cores closest to its GPU (NVML's ideal affinity), so that the pinned staging
    buffers of the host-resident path are first-touched on the GPU's NUMA node.  One process per GPU on a two-socket
    host otherwise streams half of its H2D traffic across the socket link.  Returns the CPU set, or None when NVML
    is unavailable (the binding is an optimisation, never a requirement)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        pynvml.nvmlShutdown()
        return cpus
    except Exception:
        return None


def shard_bounds(n, rank, world, align=1):
    """Contiguous split of ``n`` items (x-rows) into ``world`` shards whose edges are multiples of ``align`` (the
    binning factor); the remainder goes to the first shards."""
    units = n // align
    base, rem = divmod(units, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    lo, hi = lo * align, hi * align
    if rank == world - 1:
        hi = n
    return lo, hi


def split_even(n, rank, world):
    """Even split of a compacted active list (fit_region with a sparse mask): load balance over ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _control_device(device):
    """Device of the small control-plane tensors: the GPU under NCCL, the CPU under gloo."""
    return device if dist.get_backend() == 'nccl' else torch.device('cpu')


class PeerRows:
    """Rows ``[n_total, K]`` float32 and status ``[n_total]`` int32 on rank ``dst``, writable by every rank.

    Collective: every rank of the process group constructs it with the same arguments.  ``dst`` allocates the two
    tensors (``.rows`` / ``.status``), the others map them (same node; works between processes that share one GPU
    as well) and get addresses.  ``window(lo, hi)`` returns what ``engine.fit_batch(rows=..., status=...)`` takes for
    the spaxels ``[lo, hi)`` of the global list.  ``complete()`` = this rank's stream drained + barrier: afterwards
    ``dst`` may read.  ``ok`` is False on EVERY rank when any rank could not map the buffers (callers then use
    :func:`gather_rows`)."""

    def __init__(self, n_total, K, device, dst=0):
        from . import _capi
        from .engine import DevPtr
        self._DevPtr = DevPtr
        self.n, self.K, self.dst, self.device = int(n_total), int(K), int(dst), torch.device(device)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.rows = self.status = None
        self._opened = []
        lib = _capi.lib()
        payload = [None]
        good = 1
        if self.rank == dst:
            self.rows = torch.empty((self.n, self.K), dtype=torch.float32, device=self.device)
            self.status = torch.empty((self.n,), dtype=torch.int32, device=self.device)
            items = []
            for t in (self.rows, self.status):
                h = (C.c_ubyte * 64)()
                off = C.c_uint64(0)
                rc = lib.lucifit_peer_export(C.c_void_p(t.data_ptr()), h, C.byref(off)) if t.numel() else 0
                if rc != 0:
                    good = 0
                items.append((bytes(h), int(off.value)))
            payload = [items]
        dist.broadcast_object_list(payload, src=dst)
        self._addr = [0, 0]
        if self.rank != dst and self.n > 0:
            with torch.cuda.device(self.device):
                for i, (h, off) in enumerate(payload[0]):
                    out = C.c_void_p(0)
                    hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    rc = lib.lucifit_peer_open(hb, C.c_uint64(off), C.byref(out))
                    if rc != 0 or not out.value:
                        good = 0
                        break
                    self._addr[i] = int(out.value)
                    self._opened.append((int(out.value), off))
        flag = torch.tensor([good], dtype=torch.int32, device=_control_device(self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(int(flag.item()) == 1)

    def window(self, lo, hi):
        if self.rank == self.dst:
            return self.rows[lo:hi], self.status[lo:hi]
        D = self._DevPtr
        return (D(self._addr[0] + 4 * self.K * int(lo), (hi - lo, self.K)), D(self._addr[1] + 4 * int(lo), (hi - lo,)))

    def complete(self):
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier()

    def close(self):
        if self._opened:
            from . import _capi
            lib = _capi.lib()
            with torch.cuda.device(self.device):
                for addr, off in self._opened:
                    lib.lucifit_peer_close(C.c_void_p(addr), C.c_uint64(off))
            self._opened = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_rows(local_rows, dst=0, counts=None, out=None):
    """Gather ``[n_r, K]`` row blocks of all ranks to ``dst`` in rank order (concatenated ``[sum n_r, K]``); other
    ranks get ``None``.  Shards may differ in length.

    ``counts`` (rows per rank, e.g. from :func:`shard_bounds`) avoids the exchange of the block lengths, which costs a
    host synchronisation; ``out`` is an optional preallocated ``[sum n_r, K]`` destination on ``dst``.  Every block
    is received straight into its slice of the destination (point-to-point, no padding, no concatenation)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    world, rank = dist.get_world_size(), dist.get_rank()
    if counts is None:
        n_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
        got = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(got, n_local)
        counts = [int(c.item()) for c in got]
    K = local_rows.shape[1]
    staged = local_rows.is_cuda and dist.get_backend() != 'nccl'     # gloo moves host memory: stage device rows through it
    if rank != dst:
        if local_rows.shape[0] > 0:
            dist.send(local_rows.cpu() if staged else local_rows.contiguous(), dst=dst)
        return None
    total = sum(counts)
    if out is None:
        out = torch.empty((total, K), dtype=local_rows.dtype, device=local_rows.device)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    reqs = []
    for r in range(world):
        if counts[r] == 0:
            continue
        blk = out[offs[r]:offs[r + 1]]
        if r == dst:
            blk.copy_(local_rows)
        elif staged:
            tmp = torch.empty(blk.shape, dtype=blk.dtype)
            reqs.append((dist.irecv(tmp, src=r), blk, tmp))
        else:
            reqs.append((dist.irecv(blk, src=r), None, None))
    for q, blk, tmp in reqs:
        q.wait()
        if tmp is not None:
            blk.copy_(tmp)
    return out


def max_over_ranks(value, device=None):
    """max of a python float over ranks (timing)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
