"""End-to-end entry for HOST buffers: keymask discovery of a list of videos whose label maps, tracks and visibility
flags live in (page-locked) host memory - what a caller has when the upstream producers ran elsewhere. Chunks of
videos are staged into two device sets on two copy streams, so the host -> device DMA of chunk c + 1 overlaps the
kernels of chunk c; per chunk only the small result tables travel back (status, clusters, windows, group labels,
one2x flags: ~70 KB per video).

This is the path bench.py's `e2e` record times. It is bound by the DMA (61 GB per C2 batch at ~55 GB/s per GPU
against 14 ms of kernels), which is why the wire format matters: visibility flags travel bit-packed
(S2D_DESC_VIS_BITS, 1/8 of the bytes), tracks stay float32 as the tracker produced them."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from .pipeline import Batch, Params, VideoInput


@dataclass
class HostVideo:
    """host tensors of one video (ideally inside a hostmem.PinnedPool)"""
    labels: torch.Tensor                 # u8 [T,H,W]
    tracks: torch.Tensor                 # f32 [Nm,T,P,2]
    vis: torch.Tensor                    # u8 [Nm,T,P] or int32 [Nm,T,ceil(P/32)] when vis_bits
    vis_bits: bool = False
    max_label: Optional[int] = None

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.labels, self.tracks, self.vis))

    def signature(self):
        return (tuple(self.labels.shape), tuple(self.tracks.shape), tuple(self.vis.shape), self.vis_bits, self.max_label)


class HostPipeline:
    """Two staging sets + batches for chunks of `chunk` videos. `like` gives the shapes of the slots: `chunk` entries
    (both sets alike) or 2 * `chunk` entries (even chunks land in the first half's shapes, odd chunks in the second
    half's). The allocation happens here, once; run() only enqueues copies and kernels."""

    RESULT_TABLES = ("vidinfo", "clusterinfo", "rowinfo", "glabel", "one2x")

    def __init__(self, like: Sequence[HostVideo], device, params: Params = Params(), chunk: Optional[int] = None):
        self.device = torch.device(device)
        self.params = params
        self.chunk = int(chunk) if chunk else len(like)
        assert len(like) in (self.chunk, 2 * self.chunk)
        halves = [list(like[:self.chunk]), list(like[-self.chunk:])]
        self.sigs = [[v.signature() for v in h] for h in halves]
        self.sets = []
        with torch.cuda.device(self.device):
            for h in halves:
                dv = [VideoInput(torch.empty(v.labels.shape, dtype=torch.uint8, device=self.device),
                                 torch.empty(v.tracks.shape, dtype=torch.float32, device=self.device),
                                 torch.empty(v.vis.shape, dtype=v.vis.dtype, device=self.device),
                                 vis_bits=v.vis_bits, max_label=v.max_label) for v in h]
                b = Batch(dv, device=self.device)
                self.sets.append((dv, b))
            self.copy_streams = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
            self.compute_stream = torch.cuda.Stream(self.device)
        self.host_results: List[dict] = []          # one set of pinned result tables per chunk, grown on first use
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0

    def run(self, videos: Sequence[HostVideo]) -> List[dict]:
        """All videos, chunk by chunk. Returns one dict of pinned host result tables per chunk (owned by the pipeline,
        overwritten by the next run). Synchronises the device once, at the end."""
        assert len(videos) % self.chunk == 0, "the list must be a whole number of chunks"
        nchunks = len(videos) // self.chunk
        while len(self.host_results) < nchunks:      # first call only: pinned result tables for every chunk
            b0 = self.sets[len(self.host_results) % 2][1]
            self.host_results.append({k: torch.empty(getattr(b0, k).shape, dtype=getattr(b0, k).dtype, pin_memory=True)
                                      for k in self.RESULT_TABLES})
        done = [None, None]
        self.h2d_bytes = self.d2h_bytes = self.launches = 0
        with torch.cuda.device(self.device):
            for c in range(nchunks):
                s = c % 2
                dv, b = self.sets[s]
                if done[s] is not None:
                    self.copy_streams[s].wait_event(done[s])          # staging set s is free again
                with torch.cuda.stream(self.copy_streams[s]):
                    for j in range(self.chunk):
                        hv = videos[c * self.chunk + j]
                        assert hv.signature() == self.sigs[s][j], "video does not fit the staging slot it lands in"
                        dv[j].labels.copy_(hv.labels, non_blocking=True)
                        dv[j].tracks.copy_(hv.tracks, non_blocking=True)
                        dv[j].vis.copy_(hv.vis, non_blocking=True)
                        self.h2d_bytes += hv.nbytes
                    ready = torch.cuda.Event()
                    ready.record()
                self.compute_stream.wait_event(ready)
                with torch.cuda.stream(self.compute_stream):
                    self.launches += b.run(self.params)
                    for k, hbuf in self.host_results[c].items():
                        hbuf.copy_(getattr(b, k), non_blocking=True)
                        self.d2h_bytes += hbuf.numel() * hbuf.element_size()
                    done[s] = torch.cuda.Event()
                    done[s].record()
            self.compute_stream.synchronize()
        return self.host_results[:nchunks]
