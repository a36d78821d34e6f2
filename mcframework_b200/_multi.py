"""Multi-GPU plumbing (one process per GPU, ``torch.distributed`` / NCCL over NVLink).

The replication loop needs no data-path collective: ranks replay disjoint ranges of the spawned
blocks.  Collectives appear only where partial statistics are combined (``_stats_device``) and
here, where every rank receives the full result vector the public API returns.
"""
from __future__ import annotations


def gather_results(local, group):
    """Concatenate the ranks' result shards in rank (= reference block) order on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty(width, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


def gather_results_to_root(local, group):
    """Rank 0 receives the concatenation (rank order); other ranks get ``None``."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty(width, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


# ------------------------------------------------------------------------------------------------------------------
# Shared-memory gather (opt-in: ``sim.gather_mode = "shm"``; host logic covered by the gloo tests, the pinned copies not
# yet timed on hardware).  ``gather_results_to_root`` funnels the whole result vector through ONE GPU's PCIe link (8 GPUs:
# 6.4 GB, ~130 ms per 1e8-path shard set).  Here rank 0 owns a POSIX shared-memory segment, every rank maps it and copies
# its own shard device->host into its slice over its own link, in parallel; no data-path collective, two tiny ones for
# the bookkeeping.  Segments are pooled: a segment goes back to the pool when the ndarray handed to the caller is garbage
# collected, so steady-state runs reuse mappings that are already page-locked (``cudaHostRegister`` is paid once).
# ------------------------------------------------------------------------------------------------------------------
class _Segment:
    __slots__ = ("shm", "nbytes", "registered", "owner", "ranges")

    def __init__(self, shm, nbytes, owner):
        self.shm, self.nbytes, self.registered, self.owner = shm, nbytes, False, owner
        self.ranges = {}  # (page-aligned offset, length) -> registered with CUDA


class _Attachment:
    """A plain mapping of somebody else's segment.  ``SharedMemory(name=...)`` would also register the name with this
    process's resource tracker (Python < 3.13), which then unlinks the owner's segment when this process exits."""

    def __init__(self, name: str, nbytes: int):
        import mmap
        import os

        import _posixshmem

        self.name = name
        fd = _posixshmem.shm_open("/" + name.lstrip("/"), os.O_RDWR, mode=0o600)
        try:
            self._mmap = mmap.mmap(fd, nbytes)
        finally:
            os.close(fd)
        self.buf = memoryview(self._mmap)

    def close(self) -> None:
        self.buf.release()
        self._mmap.close()


class SharedResultPool:
    """Per-process cache of shared-memory result segments (rank 0 creates and recycles, the others attach by name)."""

    def __init__(self):
        self._free: list[_Segment] = []    # rank 0: segments whose result array is gone
        self._attached: dict[str, _Segment] = {}  # other ranks: name -> mapping
        self._all: list[_Segment] = []
        import atexit

        atexit.register(self.close)

    # -- rank 0 ---------------------------------------------------------------------------------------------------
    def acquire(self, nbytes: int) -> _Segment:
        from multiprocessing import shared_memory

        best = None
        for seg in self._free:
            if seg.nbytes >= nbytes and (best is None or seg.nbytes < best.nbytes):
                best = seg
        if best is not None:
            self._free.remove(best)
            return best
        seg = _Segment(shared_memory.SharedMemory(create=True, size=max(int(nbytes), 8)), max(int(nbytes), 8), True)
        self._all.append(seg)
        return seg

    def release(self, seg: _Segment) -> None:
        self._free.append(seg)

    # -- other ranks ----------------------------------------------------------------------------------------------
    def attach(self, name: str, nbytes: int) -> _Segment:
        seg = self._attached.get(name)
        if seg is None:
            seg = _Segment(_Attachment(name, int(nbytes)), int(nbytes), False)
            self._attached[name] = seg
            self._all.append(seg)
        return seg

    @staticmethod
    def page_lock(seg: _Segment) -> None:
        """Register the mapping with CUDA once, so device->host copies into it run at pinned-memory speed."""
        if seg.registered:
            return
        import ctypes

        import torch

        if torch.cuda.is_available():
            addr = ctypes.addressof(ctypes.c_char.from_buffer(seg.shm.buf))
            rc = torch.cuda.cudart().cudaHostRegister(addr, seg.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed with error {int(rc)}")
        seg.registered = True

    @staticmethod
    def page_lock_range(seg: _Segment, offset: int, nbytes: int) -> bool:
        """Register only ``[offset, offset + nbytes)`` of the mapping (rounded out to pages): a rank copies into ITS slice
        of the shared result, and eight processes each pinning a whole multi-GB segment run into the locked-memory limit
        (cudaHostRegister error 304 at 8 x 6.4 GB).  Returns False when the range could not be pinned (the copy then
        runs through pageable memory: slower, still correct)."""
        import ctypes

        import torch

        if nbytes <= 0 or not torch.cuda.is_available():
            return True
        page = 4096
        lo = (offset // page) * page
        hi = min(seg.nbytes, -(-(offset + nbytes) // page) * page)
        if seg.registered or any(a <= lo and hi <= a + n for (a, n) in seg.ranges):
            return True
        base = ctypes.addressof(ctypes.c_char.from_buffer(seg.shm.buf))
        rc = int(torch.cuda.cudart().cudaHostRegister(base + lo, hi - lo, 0))
        if rc != 0:
            return False
        seg.ranges[(lo, hi - lo)] = True
        return True

    def close(self) -> None:
        import ctypes

        for seg in self._all:
            try:
                for (lo, _n) in list(seg.ranges):
                    import torch

                    if torch.cuda.is_available():
                        base = ctypes.addressof(ctypes.c_char.from_buffer(seg.shm.buf))
                        torch.cuda.cudart().cudaHostUnregister(base + lo)
                seg.ranges = {}
                if seg.registered:
                    import torch

                    if torch.cuda.is_available():
                        addr = ctypes.addressof(ctypes.c_char.from_buffer(seg.shm.buf))
                        torch.cuda.cudart().cudaHostUnregister(addr)
                seg.shm.close()
                if seg.owner:
                    seg.shm.unlink()
            except Exception:  # noqa: BLE001 - interpreter shutdown: nothing useful to do
                pass
        self._all, self._free, self._attached = [], [], {}


_POOL = None


def _pool() -> SharedResultPool:
    global _POOL
    if _POOL is None:
        _POOL = SharedResultPool()
    return _POOL


_ONE_HOST: dict = {}


def _require_one_host(group) -> None:
    """POSIX shared memory only reaches the ranks of ONE host: checked once per group, and EVERY rank raises when the
    group spans hosts (a rank that failed alone in ``shm_open`` would leave the others waiting at the barrier)."""
    import socket

    import torch.distributed as dist

    key = id(group)
    if key not in _ONE_HOST:
        names = [None] * dist.get_world_size(group)
        dist.all_gather_object(names, socket.gethostname(), group=group)
        _ONE_HOST[key] = len(set(names)) == 1
    if not _ONE_HOST[key]:
        raise RuntimeError('gather_mode="shm" needs every rank of the group on one host; use "root" or "all"')


def shared_result_buffer(sizes, group):
    """One float64 result buffer of ``sum(sizes)`` elements in POSIX shared memory, mapped and page-locked by every rank
    of the (single-host) group.  Returns ``(whole ndarray, this rank's slice as a torch tensor, segment)``; rank 0 owns the
    segment -- it goes back to the pool when rank 0's ndarray is garbage collected (``finish_shared_result``)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    _require_one_host(group)
    total, offset = sum(sizes), sum(sizes[:rank])
    pool = _pool()
    seg = pool.acquire(total * 8) if rank == 0 else None
    box = [(seg.shm.name, seg.nbytes) if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        seg = pool.attach(*box[0])
    pool.page_lock_range(seg, offset * 8, sizes[rank] * 8)  # only the slice this rank's GPU writes
    whole = np.ndarray((total,), dtype=np.float64, buffer=seg.shm.buf)
    mine = torch.from_numpy(whole[offset:offset + sizes[rank]])
    return whole, mine, seg


def finish_shared_result(whole, seg, sizes, group):
    """Barrier (every shard has landed), then rank 0 gets the whole vector and the other ranks a view of their own shard
    (valid until rank 0 drops the result and the segment is reused by a later run)."""
    import weakref

    import torch.distributed as dist

    rank = dist.get_rank(group)
    dist.barrier(group=group)
    if rank != 0:
        offset = sum(sizes[:rank])
        return whole[offset:offset + sizes[rank]]
    weakref.finalize(whole, _pool().release, seg)  # the segment is reusable once the caller has dropped the result
    return whole


def gather_results_shared(local, group):
    """Rank 0 returns the concatenation (rank order) as a host float64 ndarray backed by shared memory; the other ranks
    return ``None``.  Every rank copies only its own shard, straight from its device into the shared buffer."""
    import weakref

    import numpy as np
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    _require_one_host(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    total, offset = sum(sizes), sum(sizes[:rank])
    pool = _pool()
    seg = pool.acquire(total * 8) if rank == 0 else None
    box = [(seg.shm.name, seg.nbytes) if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        seg = pool.attach(*box[0])
    pool.page_lock_range(seg, offset * 8, sizes[rank] * 8)
    whole = np.ndarray((total,), dtype=np.float64, buffer=seg.shm.buf)
    if sizes[rank]:
        torch.from_numpy(whole[offset:offset + sizes[rank]]).copy_(local.to(torch.float64), non_blocking=False)
    dist.barrier(group=group)  # every shard has landed
    if rank != 0:
        del whole
        return None
    weakref.finalize(whole, pool.release, seg)  # the segment is reusable once the caller has dropped the result
    return whole
