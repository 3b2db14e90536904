"""Host side of the end-to-end path: where the staging buffers live and which cores feed them.

The device-resident path runs at TB/s; fed from host buffers it is bound by the host -> device DMA (61 GB per C2
batch). Two things decide what that DMA reaches, above all with several GPUs in one box:

  * placement - the pinned pool of a rank must sit on the NUMA node its GPU hangs off, and the rank's threads must
    run there (first-touch allocation follows the thread), otherwise every byte crosses the socket interconnect;
  * page size - a pool of 4 KB pages costs one IOMMU translation per 4 KB of DMA (in a VM with passed-through GPUs
    that is the bottleneck); the pool is therefore an anonymous mapping madvise'd to transparent huge pages, touched
    once, and page-locked with cudaHostRegister (s2d_host_register) instead of coming from cudaHostAlloc.

torch is plumbing here: tensors are views of the mapping, copies are torch's non_blocking copy_ (cudaMemcpyAsync)."""
from __future__ import annotations

import ctypes as C
import mmap
import os
from typing import List, Optional

import numpy as np
import torch

from . import _lib

MADV_HUGEPAGE = 14
_HUGE = 2 << 20


def _parse_cpulist(txt: str) -> List[int]:
    cpus: List[int] = []
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device from sysfs, None when the platform does not say (VMs report -1)."""
    buf = C.create_string_buffer(32)
    try:
        _lib.call("s2d_device_pci_bus_id", int(device_index), buf, 32)
        path = f"/sys/bus/pci/devices/{buf.value.decode().lower()}/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError, _lib.S2DError):
        return None


def bind_to_gpu(device_index: int, local_rank: int = 0, local_world: int = 1) -> dict:
    """Pin the calling process (all its current threads' future children) to the cores next to `device_index`: the
    GPU's NUMA node when sysfs knows it, otherwise an even slice of the allowed cores per local rank (on the usual
    two-socket boxes GPUs 0-3 / 4-7 sit on sockets 0 / 1, so the slice order follows the device order). Call it
    BEFORE allocating pinned pools. Returns what was done (for the bench record)."""
    allowed = sorted(os.sched_getaffinity(0))
    info = {"numa_node": None, "cpus": len(allowed), "how": "unchanged"}
    node = gpu_numa_node(device_index)
    cpus: List[int] = []
    if node is not None:
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = [c for c in _parse_cpulist(f.read()) if c in allowed]
            info.update(numa_node=node, how="numa node of the GPU (sysfs)")
        except OSError:
            cpus = []
    if not cpus and local_world > 1 and len(allowed) >= local_world:
        k = len(allowed) // local_world
        cpus = allowed[local_rank * k:(local_rank + 1) * k]
        info.update(how=f"even slice of the allowed cores per local rank ({k} each)")
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
        except OSError as e:
            info["how"] = f"sched_setaffinity failed: {e}"
    return info


class PinnedPool:
    """One page-locked, huge-page backed host mapping carved into tensors. Lives as long as the object."""

    def __init__(self, nbytes: int, huge: bool = True):
        self.nbytes = (int(nbytes) + _HUGE - 1) // _HUGE * _HUGE
        self.map = mmap.mmap(-1, self.nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        self.buf = np.frombuffer(self.map, dtype=np.uint8)
        self.addr = self.buf.ctypes.data
        self.huge = False
        if huge:
            libc = C.CDLL(None, use_errno=True)
            libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
            self.huge = libc.madvise(C.c_void_p(self.addr), self.nbytes, MADV_HUGEPAGE) == 0
        self.buf[::4096] = 0                      # first touch: pages land on the node this thread runs on
        _lib.call("s2d_host_register", self.addr, self.nbytes)
        self.off = 0
        self.registered = True

    def take(self, shape, dtype: torch.dtype) -> torch.Tensor:
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = (self.off + 255) // 256 * 256
        if off + n > self.nbytes:
            raise MemoryError(f"PinnedPool: {n} bytes requested, {self.nbytes - off} left")
        self.off = off + n
        t = torch.from_numpy(self.buf[off:off + n]).view(dtype).reshape(tuple(shape))
        return t

    def close(self):
        if self.registered:
            _lib.call("s2d_host_unregister", self.addr)
            self.registered = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pooled_copies(tensors: List[torch.Tensor], huge: bool = True):
    """Host copies of `tensors` (any device) inside one PinnedPool; returns (pool, [host tensors])."""
    total = sum((t.numel() * t.element_size() + 255) // 256 * 256 for t in tensors) + 4096
    pool = PinnedPool(total, huge=huge)
    out = []
    for t in tensors:
        h = pool.take(t.shape, t.dtype)
        h.copy_(t)
        out.append(h)
    return pool, out
