"""Seeded synthetic videos for the keymask-discovery hot path.

A *scene* is what the reference's per-video loop consumes once the upstream producers
(CutS3D masks, CoTracker) have run:

  labels  u8  [T, H, W]      per-frame label maps, 0 = background, ids are per-frame
                             ranks exactly as load_masks() builds them
                             (keymask_ident/cotracker_matching.py:54-71)
  tracks  f32 [Nm, T, P, 2]  CoTracker pred_tracks (x, y) of the P points sampled inside mask q,
                             one row block per (frame, mask) query in global-id order
                             (cotracker_matching.py:289-306)
  vis     u8  [Nm, T, P]     CoTracker pred_visibility flags of the same points
                             (cotracker_occlusions.py:355-359)

The host generator (numpy) is used by tests, golden generation and the CPU baseline; the
device generator (torch, plumbing only - it is outside every timed region) builds the large
BASELINE.json configurations directly in HBM.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Scene:
    labels: np.ndarray            # u8 [T,H,W]
    tracks: np.ndarray            # f32 [Nm,T,P,2]
    vis: np.ndarray               # u8 [Nm,T,P]
    query_frame: np.ndarray       # i32 [Nm]
    query_label: np.ndarray       # i32 [Nm]
    object_of_query: np.ndarray   # i32 [Nm] (generator bookkeeping, not an input of the path)
    colors: np.ndarray = field(default=None)  # u8 [K,3] palette used when written as colour PNGs

    @property
    def shape(self):
        T, H, W = self.labels.shape
        return T, H, W, self.tracks.shape[0], self.tracks.shape[2]


def enumerate_queries(labels: np.ndarray):
    """(frame, label) pairs in global-id order: per frame the sorted unique labels minus the
    smallest one (cotracker_matching.py:294 - the [1:] drops the first unique value even when
    it is not background)."""
    qf, ql = [], []
    for t in range(labels.shape[0]):
        ids = np.unique(labels[t])[1:]
        for o in ids:
            qf.append(t)
            ql.append(int(o))
    return np.asarray(qf, np.int32), np.asarray(ql, np.int32)


def _paint(T, H, W, objs, rng, occluded_obj, occl_span, full_cover_frames):
    """Paint object index maps [T,H,W] (int16, -1 background) with later objects on top."""
    yy, xx = np.mgrid[0:H, 0:W]
    omap = np.full((T, H, W), -1, np.int16)
    for t in range(T):
        if t in full_cover_frames:
            # no background pixel in this frame: vertical stripes of the objects
            k = len(objs)
            omap[t] = (xx * k // W).astype(np.int16)
            continue
        for k, o in enumerate(objs):
            if k == occluded_obj and occl_span[0] <= t < occl_span[1]:
                continue
            cx = o["cx"] + o["vx"] * t
            cy = o["cy"] + o["vy"] * t
            if o["ellipse"]:
                m = ((xx - cx) / o["rx"]) ** 2 + ((yy - cy) / o["ry"]) ** 2 <= 1.0
            else:
                m = (np.abs(xx - cx) <= o["rx"]) & (np.abs(yy - cy) <= o["ry"])
            omap[t][m] = k
    return omap


def make_scene(seed: int, T: int, H: int, W: int, M: int, P: int, *, occlude: bool = True,
               full_cover_frames=(), specials: bool = False, noise: float = 0.7,
               dup_rate: float = 0.0) -> Scene:
    """M moving rectangles/ellipses; object 0 is absent for the middle third of the video so
    that at least two visibility clusters appear. `specials` injects NaN/inf, half-integer and
    out-of-range coordinates; `dup_rate` forces exact duplicate points."""
    rng = np.random.default_rng(seed)
    objs = []
    for k in range(M):
        rx = rng.uniform(0.05, 0.12) * W
        ry = rng.uniform(0.06, 0.14) * H
        objs.append(dict(cx=rng.uniform(rx, W - rx), cy=rng.uniform(ry, H - ry), rx=rx, ry=ry,
                         vx=rng.uniform(-0.006, 0.006) * W, vy=rng.uniform(-0.006, 0.006) * H,
                         ellipse=bool(k % 2)))
    occl_span = (T // 3, (2 * T) // 3) if occlude else (0, 0)
    omap = _paint(T, H, W, objs, rng, 0 if occlude else -1, occl_span, set(full_cover_frames))

    # per-frame rank labels (what load_masks produces from the colour PNGs)
    labels = np.zeros((T, H, W), np.uint8)
    present = []
    for t in range(T):
        ks = np.unique(omap[t])
        ks = ks[ks >= 0]
        present.append(ks)
        lut = np.zeros(M + 1, np.uint8)
        lut[ks + 1] = np.arange(1, len(ks) + 1, dtype=np.uint8)
        labels[t] = lut[omap[t].astype(np.int32) + 1]
    area = np.zeros((T, M), np.int64)
    for t in range(T):
        for k in present[t]:
            area[t, k] = int((omap[t] == k).sum())

    qf, ql = enumerate_queries(labels)
    Nm = len(qf)
    tracks = np.empty((Nm, T, P, 2), np.float32)
    vis = np.empty((Nm, T, P), np.uint8)
    obj_of_q = np.empty(Nm, np.int32)
    dts = np.arange(T, dtype=np.float64)
    for q in range(Nm):
        t, lab = int(qf[q]), int(ql[q])
        ys, xs = np.nonzero(labels[t] == lab)
        k = int(omap[t][ys[0], xs[0]])
        obj_of_q[q] = k
        sel = rng.integers(0, len(ys), size=P)
        if dup_rate > 0:
            ndup = int(P * dup_rate)
            sel[P - ndup:] = sel[:ndup]
        jit = rng.uniform(-0.4, 0.4, size=(P, 2)) if dup_rate == 0 else np.zeros((P, 2))
        px = xs[sel] + jit[:, 0]
        py = ys[sel] + jit[:, 1]
        o = objs[k]
        dx = o["vx"] * (dts - t)
        dy = o["vy"] * (dts - t)
        nz = rng.normal(0.0, noise, size=(T, P, 2)) if noise > 0 else np.zeros((T, P, 2))
        nz[t] = 0.0
        tracks[q, :, :, 0] = (px[None, :] + dx[:, None] + nz[..., 0]).astype(np.float32)
        tracks[q, :, :, 1] = (py[None, :] + dy[:, None] + nz[..., 1]).astype(np.float32)
        pvis = np.where(area[:, k] > 0, 0.95, 0.05)
        vis[q] = (rng.random((T, P)) < pvis[:, None]).astype(np.uint8)
    if specials and Nm > 0:
        # NaN/inf -> INT64_MIN -> dropped; x.5 -> round-half-even; W-0.5 -> W -> dropped
        # (cotracker_matching.py:472-479)
        n = max(1, P // 16)
        for q in range(Nm):
            tt = rng.integers(0, T, size=n)
            pp = rng.integers(0, P, size=n)
            vals = rng.choice(np.asarray([np.nan, np.inf, -np.inf, W - 0.5, -0.5, 0.5, 1.5, 2.5,
                                          -0.49, 1e20, -3e9, H - 0.5], np.float32), size=n)
            cc = rng.integers(0, 2, size=n)
            tracks[q, tt, pp, cc] = vals
    palette = np.zeros((M, 3), np.uint8)
    for k in range(M):  # strictly increasing lexicographically -> rank order == object order
        palette[k] = (10 + 2 * k, (37 * k + 11) % 256, (91 * k + 5) % 256)
    return Scene(labels, tracks, vis, qf, ql, obj_of_q, palette)


def scene_object_map(scene: Scene) -> np.ndarray:
    """Recover per-pixel palette indices [T,H,W] (-1 background) for writing colour PNGs."""
    T, H, W = scene.labels.shape
    omap = np.full((T, H, W), -1, np.int16)
    for q in range(len(scene.query_frame)):
        t, lab = int(scene.query_frame[q]), int(scene.query_label[q])
        omap[t][scene.labels[t] == lab] = scene.object_of_query[q]
    # the dropped (smallest) label of frames without background is not a query: recover it
    for t in range(T):
        ids = np.unique(scene.labels[t])
        if ids[0] != 0:
            # stripes frame: object index == label-1 by construction
            omap[t][scene.labels[t] == ids[0]] = int(ids[0]) - 1
    return omap


# --------------------------------------------------------------------------------------
# Device-side generator for the BASELINE.json-sized configurations (torch = plumbing only)
# --------------------------------------------------------------------------------------

def make_scene_device(seed: int, T: int, H: int, W: int, M: int, P: int, device, *,
                      noise: float = 0.7, occlude: bool = True, point_order: str = "raster",
                      window: int = 0, vis_bits: bool = False):
    """Same scene family as make_scene, built directly in HBM with torch ops (vectorised per
    frame). Returns dict(labels u8 [T,H,W], tracks f32 [Nm,T,P,2], vis u8 [Nm,T,P],
    query_frame i32 [Nm], query_label i32 [Nm]). Background is always present.

    `window` > 0 (long videos): full-length tracks are never materialised (SA-V-shaped videos would need 177 GB
    each). The dict then holds `tracks` = an UNINITIALISED f32 [Nm, window, P, 2] buffer, `tstart` i32 [Nm] (zeros)
    and `fill_window(tstart)`, which writes the tracks of frames [tstart[q], tstart[q] + window) of every query into
    the buffer - what a windowed tracker run hands over once stage B has produced the windows.
    `vis_bits`: the flags come bit-packed (int32 [Nm,T,ceil(P/32)], S2D_DESC_VIS_BITS); they are drawn word-wise
    (OR / AND of four random words: 15/16 visible while the object is present, 1/16 while it is not) so that a
    300-frame video's 22 G flags never exist as bytes."""
    import torch

    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    u = lambda lo, hi, n: torch.rand(n, generator=g, dtype=torch.float64) * (hi - lo) + lo
    rx = u(0.05, 0.12, M) * W
    ry = u(0.06, 0.14, M) * H
    cx = rx + torch.rand(M, generator=g, dtype=torch.float64) * (W - 2 * rx)
    cy = ry + torch.rand(M, generator=g, dtype=torch.float64) * (H - 2 * ry)
    vscale = min(1.0, 36.0 / T)                       # long videos: same total displacement as a 36-frame one
    vx = u(-0.006, 0.006, M) * W * vscale
    vy = u(-0.006, 0.006, M) * H * vscale
    occl = (T // 3, (2 * T) // 3) if occlude else (0, 0)

    f32 = torch.float32
    yy = torch.arange(H, device=device, dtype=f32)[None, :, None]
    xx = torch.arange(W, device=device, dtype=f32)[None, None, :]
    labels = torch.empty((T, H, W), dtype=torch.uint8, device=device)
    area = torch.zeros((T, M + 1), dtype=torch.int64, device=device)
    pres_l, rank_l = [], []
    TC = max(1, min(T, (1 << 26) // (H * W)))         # frames painted at a time (bounded temporaries)
    for t0 in range(0, T, TC):
        t1 = min(T, t0 + TC)
        ts = torch.arange(t0, t1, device=device, dtype=f32)[:, None, None]
        omap = torch.full((t1 - t0, H, W), -1, dtype=torch.int16, device=device)
        for k in range(M):
            ccx = float(cx[k]) + float(vx[k]) * ts
            ccy = float(cy[k]) + float(vy[k]) * ts
            if k % 2:
                m = ((xx - ccx) / float(rx[k])) ** 2 + ((yy - ccy) / float(ry[k])) ** 2 <= 1.0
            else:
                m = ((xx - ccx).abs() <= float(rx[k])) & ((yy - ccy).abs() <= float(ry[k]))
            if k == 0 and occl[1] > occl[0]:
                lo, hi = max(occl[0], t0) - t0, min(occl[1], t1) - t0
                if hi > lo:
                    m[lo:hi] = False
            omap[m] = k
            del m
        flat = omap.reshape(t1 - t0, -1).long() + 1
        a = torch.zeros((t1 - t0, M + 1), dtype=torch.int64, device=device)
        a.scatter_add_(1, flat, torch.ones_like(flat))
        area[t0:t1] = a
        pres = a[:, 1:] > 0
        rank = torch.cumsum(pres.long(), dim=1) * pres.long()      # 1-based rank among present
        lut = torch.cat([torch.zeros((t1 - t0, 1), dtype=torch.long, device=device), rank], dim=1)
        labels[t0:t1] = torch.gather(lut, 1, flat).reshape(t1 - t0, H, W).to(torch.uint8)
        del flat, omap
    pres = area[:, 1:] > 0                                     # [T,M]

    pres_h = pres.cpu()
    n_t = pres_h.sum(1).tolist()
    Nm = int(sum(n_t))
    qf = torch.repeat_interleave(torch.arange(T), torch.tensor(n_t)).to(torch.int32)
    ql = torch.cat([torch.arange(1, n + 1) for n in n_t]).to(torch.int32)

    dg = torch.Generator(device=device)
    dg.manual_seed(seed * 7919 + 13)
    Ttr = window if window > 0 else T
    tracks = torch.empty((Nm, Ttr, P, 2), dtype=f32, device=device)
    PW = (P + 31) // 32
    vis = torch.empty((Nm, T, PW), dtype=torch.int32, device=device) if vis_bits else \
        torch.empty((Nm, T, P), dtype=torch.uint8, device=device)
    vx_d = vx.to(device=device, dtype=f32)
    vy_d = vy.to(device=device, dtype=f32)
    dts = torch.arange(T, device=device, dtype=f32)
    pvis_obj = torch.where(pres, 0.95, 0.05).to(f32)            # [T,M]
    base = torch.empty((Nm, P, 2), dtype=f32, device=device) if window > 0 else None
    vel = torch.empty((Nm, 2), dtype=f32, device=device) if window > 0 else None
    row = 0
    for t in range(T):
        n = n_t[t]
        if n == 0:
            continue
        ks = torch.nonzero(pres[t]).squeeze(1)                  # object index per label 1..n
        lab_flat = labels[t].reshape(-1)
        order = torch.argsort(lab_flat, stable=True)
        counts = torch.bincount(lab_flat.long(), minlength=n + 1)
        starts = torch.cumsum(counts, 0) - counts
        r = torch.rand((n, P), generator=dg, device=device)
        if point_order == "raster":
            r = (torch.arange(P, device=device, dtype=f32)[None, :] + r) / P
        idx = starts[1:n + 1, None] + (r * counts[1:n + 1, None]).long().clamp_(max=int(counts.max()) - 1)
        idx = torch.minimum(idx, (starts[1:n + 1] + counts[1:n + 1] - 1)[:, None])
        pix = order[idx]                                        # [n,P]
        px = (pix % W).to(f32) + (torch.rand((n, P), generator=dg, device=device) - 0.5) * 0.8
        py = (pix // W).to(f32) + (torch.rand((n, P), generator=dg, device=device) - 0.5) * 0.8
        if window > 0:
            base[row:row + n, :, 0] = px
            base[row:row + n, :, 1] = py
            vel[row:row + n, 0] = vx_d[ks]
            vel[row:row + n, 1] = vy_d[ks]
        else:
            nz = torch.randn((n, T, P, 2), generator=dg, device=device) * noise
            nz[:, t] = 0
            dt = dts - t
            blk = tracks[row:row + n]
            blk[..., 0] = px[:, None, :] + (vx_d[ks][:, None] * dt[None, :])[:, :, None] + nz[..., 0]
            blk[..., 1] = py[:, None, :] + (vy_d[ks][:, None] * dt[None, :])[:, :, None] + nz[..., 1]
        pv = pvis_obj[:, ks].t()                                # [n,T]
        if vis_bits:
            w = torch.randint(-2 ** 31, 2 ** 31, (4, n, T, PW), generator=dg, device=device, dtype=torch.int64).to(torch.int32)
            hi_ = w[0] | w[1] | w[2] | w[3]                     # each bit set with probability 15/16
            lo_ = w[0] & w[1] & w[2] & w[3]                     # ... 1/16
            wv = torch.where((pv > 0.5)[:, :, None], hi_, lo_)
            if P % 32:
                wv[:, :, -1] &= (1 << (P % 32)) - 1
            vis[row:row + n] = wv
            del w, hi_, lo_, wv
        else:
            vis[row:row + n] = (torch.rand((n, T, P), generator=dg, device=device) < pv[:, :, None]).to(torch.uint8)
        row += n
    out = dict(labels=labels, tracks=tracks, vis=vis, query_frame=qf.to(device), query_label=ql.to(device))
    if window > 0:
        qf_d = out["query_frame"]

        def fill_window(tstart, rows_per_pass: int = 256):
            """tracks[q, j] = point + velocity * (tstart[q] + j - frame(q)) + N(0, noise), exact at the query frame"""
            ts_ = tstart.to(device=device, dtype=f32)
            j = torch.arange(Ttr, device=device, dtype=f32)
            for r0 in range(0, Nm, rows_per_pass):
                r1 = min(Nm, r0 + rows_per_pass)
                dt = ts_[r0:r1, None] + j[None, :] - qf_d[r0:r1, None].to(f32)          # [n,Ttr]
                nz = torch.randn((r1 - r0, Ttr, P, 2), generator=dg, device=device) * noise
                nz *= (dt != 0)[:, :, None, None]
                nz += base[r0:r1, None, :, :]
                nz += dt[:, :, None, None] * vel[r0:r1, None, None, :]
                tracks[r0:r1] = nz
                del nz

        out["tstart"] = torch.zeros(Nm, dtype=torch.int32, device=device)
        out["fill_window"] = fill_window
    return out
