"""Multi-GPU story of the hot path: videos are independent, so the work is partitioned by video
(the reference does it with contiguous --job-id/--videos-per-job slices, main_keymask_ident.py:
20-23) and the tiny per-video results are gathered on the host. No collective touches the data
path; torch.distributed is used only for the final gather of python result objects.

Partitioning is longest-processing-time-first over a cost estimate instead of contiguous slices:
video lengths vary by an order of magnitude across datasets."""
from __future__ import annotations

import heapq
from typing import Callable, List, Sequence


def video_cost(T: int, H: int, W: int, Nm: int, P: int) -> float:
    """bytes the dominant kernels stream for one video: tracks of every (query, frame) tile, the
    visibility flags and the label maps."""
    return 8.0 * Nm * T * P + 1.0 * Nm * T * P + 1.0 * T * H * W


def lpt_partition(costs: Sequence[float], nparts: int) -> List[List[int]]:
    """Deterministic LPT: items by decreasing cost (ties by index) onto the least-loaded part
    (ties by part id). Returns item indices per part, each ascending."""
    assert nparts >= 1
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    heap = [(0.0, p) for p in range(nparts)]
    heapq.heapify(heap)
    parts: List[List[int]] = [[] for _ in range(nparts)]
    for i in order:
        load, p = heapq.heappop(heap)
        parts[p].append(i)
        heapq.heappush(heap, (load + float(costs[i]), p))
    return [sorted(p) for p in parts]


def contiguous_partition(n: int, nparts: int) -> List[List[int]]:
    """the reference's scheme: job j takes [j*k, (j+1)*k) with k = ceil(n / nparts)."""
    k = -(-n // nparts) if n else 0
    return [list(range(j * k, min(n, (j + 1) * k))) for j in range(nparts)]


def run_partitioned(items: Sequence, costs: Sequence[float], worker: Callable[[List[int]], list],
                    rank: int = 0, world_size: int = 1, gather: bool = True):
    """Each rank runs `worker(indices)` on its share (returning one result per index); rank 0
    receives the results of all ranks in item order, so the output does not depend on the number
    of GPUs. With world_size == 1 no process group is needed."""
    parts = lpt_partition(costs, world_size)
    mine = parts[rank]
    local = worker(mine)
    assert len(local) == len(mine)
    if world_size == 1:
        return dict(zip(mine, local)) if not gather else [r for _, r in sorted(zip(mine, local))]
    import torch.distributed as dist
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(list(zip(mine, local)), gathered, dst=0)
    if rank != 0:
        return None
    merged = sorted((pair for part in gathered for pair in part), key=lambda x: x[0])
    assert [i for i, _ in merged] == list(range(len(items)))
    return [r for _, r in merged]
