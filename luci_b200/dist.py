"""Sharding of the spaxel list over GPUs of one node and the final gather of the output maps.

The fit path has no exchange step (every spaxel is independent; the reference's unit of parallelism is a y-row
handed to a joblib worker, ``LuciBase.py:514-517``).  Each rank owns a contiguous slab of x-rows of the cube, fits
it with no collective, and only the ``(7L+5)`` float32 output columns are gathered at the end -- over NCCL when
the tensors are on the GPU, over gloo in the CPU tests.
"""
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
    if rank != dst:
        if local_rows.shape[0] > 0:
            dist.send(local_rows.contiguous(), dst=dst)
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
        else:
            reqs.append(dist.irecv(blk, src=r))
    for q in reqs:
        q.wait()
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
