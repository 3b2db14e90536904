"""TEST INFRASTRUCTURE / CPU BASELINE ONLY - dense torch-CPU port of the reference's stage-D
arithmetic, written to follow the reference's *cost profile* (full-frame H x W tensor ops per
(query, frame, mask) pair), not the sparse shortcut of keymask_oracle.point_votes.

This is what `bench.py` times on the GPU box's host cores as `cpu_baseline` (kind "port") and
under `--impl reference`: the reference itself is Python and lives only in the build container
(/root/reference does not exist on the GPU box), so its algorithm is restated here step by step:

  rasterise tracks              cotracker_matching.py:453-503 (return_mask=False)
  per frame: unique labels      cotracker_matching.py:682-683
  per mask: extract, resize     cotracker_matching.py:687-689 (get_segmentation_mask :176-209)
  AND / OR / sum                cotracker_matching.py:640-662
  threshold + one-to-many       cotracker_matching.py:710-717, 1082-1111
  visibility mean               cotracker_occlusions.py:359

Checked against the reference goldens in tests/test_oracle_golden.py (test_dense_port_*).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def track_rasters(tracks_q: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """[T,P,2] float32 -> [T,H,W] uint8 rasters of the rounded in-bounds points."""
    T = tracks_q.shape[0]
    out = torch.zeros((T, H, W), dtype=torch.uint8)
    pts = tracks_q.round().long()
    for t in range(T):
        x, y = pts[t, :, 0], pts[t, :, 1]
        keep = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        sel = pts[t][keep].numpy()
        plane = np.zeros((H, W), dtype=np.uint8)
        plane[sel[:, 1], sel[:, 0]] = 1
        out[t] = torch.from_numpy(plane)
    return out


def pair_counts(raster: torch.Tensor, mask255: torch.Tensor):
    """(intersection, union) of a point raster with a 0/255 mask restricted to the raster."""
    pm = raster.bool()
    mk = mask255.bool() * pm
    return int(torch.sum(pm & mk).item()), int(torch.sum(pm | mk).item())


def match_query_dense(labels_thw1: torch.Tensor, tracks_q: torch.Tensor, v0: int, v1: int,
                      H: int, W: int, matching_threshold: float = 0.5):
    """One stage-D query: returns (matches [(frame, label)], comps [(frame, label, I, U, iou)], one2x)."""
    rasters = track_rasters(tracks_q, H, W)
    matches, comps = [], []
    multi = {}
    for t in range(v0, v1 + 1):
        frame = labels_thw1[t]
        ids = torch.sort(torch.unique(frame[..., 0])[1:])[0]
        for oid in ids:
            m = (frame[..., 0] == oid).to(torch.uint8) * 255
            m = F.interpolate(m[None, None].float(), size=(H, W), mode="nearest").to(torch.uint8)[0, 0]
            I, U = pair_counts(rasters[t], m)
            iou = 0.0 if U == 0 else I / U
            comps.append((t, int(oid), I, U, iou))
            if iou > matching_threshold:
                matches.append((t, int(oid)))
            if iou > 0.25:
                multi[t] = multi.get(t, 0) + 1
    one2x = 1 if sum(1 for c in multi.values() if c > 1) >= 5 else 0
    return matches, comps, one2x


def visibility_rows(vis: torch.Tensor) -> torch.Tensor:
    """[Nm,T,P] bool -> [Nm,T] float32 (mean over points)."""
    return torch.mean(vis.float(), dim=2)


def time_sample(labels: np.ndarray, tracks: np.ndarray, vis: np.ndarray, queries, window,
                matching_threshold: float = 0.5):
    """Run the dense port for the given query rows over `window`=(v0,v1); returns
    (seconds, n_pairs, results). Used by bench.py's cpu_baseline / --impl reference."""
    import time
    T, H, W = labels.shape
    lab = torch.from_numpy(labels.astype(np.int64))[..., None]
    t0 = time.perf_counter()
    _ = visibility_rows(torch.from_numpy(vis[queries].astype(bool)))
    out, npairs = [], 0
    for q in queries:
        m, c, o = match_query_dense(lab, torch.from_numpy(tracks[q]), window[0], window[1], H, W, matching_threshold)
        npairs += len(c)
        out.append((m, c, o))
    return time.perf_counter() - t0, npairs, out
