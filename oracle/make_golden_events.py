"""TEST INFRASTRUCTURE. Golden vectors for extract_appearance_events / boolean_visibility, produced by
the UNMODIFIED reference functions (keymask_ident/cotracker_occlusions.py:166-240; the copies in
cotracker_matching.py:212-286 are checked to agree) in the build container:

    python -m oracle.make_golden_events        # writes tests/golden/events.npz

Cases cover the default parameters, odd and even opening windows (even windows shorten the signal),
window 1, rows that start visible (the first `end` precedes the first `start`: the reference zips them
anyway), all-on / all-off rows, values exactly at the threshold, and smoothing windows > 1 on
flag-valued curves whose sums stay away from the threshold."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_functions():
    ref_harness._install_stubs(object)
    if ref_harness.REFERENCE_DIR not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_DIR)
    real_exists = os.path.exists
    os.path.exists = lambda p: True if str(p).rstrip("/") == "/mnt/data/checkpoints" else real_exists(p)
    try:
        import cotracker_occlusions
        import cotracker_matching
    finally:
        os.path.exists = real_exists
    return cotracker_occlusions, cotracker_matching


def cases():
    rng = np.random.default_rng(77)
    out = []

    def runs(n, T, p_flip):
        v = np.zeros((n, T), np.float32)
        for i in range(n):
            s = rng.integers(0, 2)
            for t in range(T):
                if rng.random() < p_flip:
                    s ^= 1
                v[i, t] = s
        return v

    # (name, V, smoothing_window, thresh, min_run_length)
    T = 36
    v = runs(24, T, 0.15) * rng.uniform(0.96, 1.0, size=(24, T)).astype(np.float32)
    out.append(("default", v, 1, 0.95, 4))
    out.append(("k3", v, 1, 0.95, 3))
    out.append(("k5", v, 1, 0.95, 5))
    out.append(("k1", v, 1, 0.95, 1))
    out.append(("k2", v, 1, 0.95, 2))
    v2 = rng.random((40, 300)).astype(np.float32)
    v2[0] = 1.0; v2[1] = 0.0; v2[2, :150] = 1.0; v2[3, 150:] = 1.0
    v2[4] = np.float32(0.3); v2[5] = np.nextafter(np.float32(0.3), np.float32(0))    # at / just below the threshold
    out.append(("long_thr03", v2, 1, 0.3, 4))
    out.append(("long_thr05_k7", v2, 1, 0.5, 7))
    flags = runs(16, 64, 0.2)
    out.append(("smooth3", flags, 3, 0.5, 4))       # sums k/3: {0, .33, .67, 1} vs 0.5
    out.append(("smooth5", flags, 5, 0.5, 3))       # sums k/5 vs 0.5
    out.append(("short", runs(5, 8, 0.3), 1, 0.95, 4))
    return out


def main():
    occ, mat = reference_functions()
    store = {}
    names = []
    for name, v, sw, th, k in cases():
        t = torch.from_numpy(v)
        ev = occ.extract_appearance_events(t, smoothing_window=sw, thresh=th, min_run_length=k)
        ev2 = mat.extract_appearance_events(t, smoothing_window=sw, thresh=th, min_run_length=k)
        assert ev == ev2, name
        n = v.shape[0]
        ns = np.asarray([len(ev[i]) for i in range(n)], np.int32)
        flat = np.asarray([p for i in range(n) for p in ev[i]], np.int32).reshape(-1, 2)
        store[f"{name}__V"] = v
        store[f"{name}__params"] = np.asarray([sw, th, k], np.float64)
        store[f"{name}__npairs"] = ns
        store[f"{name}__pairs"] = flat
        store[f"{name}__bool"] = occ.boolean_visibility(t, threshold=th).numpy()
        names.append(name)
    store["names"] = np.asarray(names)
    path = os.path.join(ROOT, "tests", "golden", "events.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, {n: int(store[f"{n}__npairs"].sum()) for n in names})

    # convex-hull grid scorers (cotracker_matching.py:506-637): random point rasters inside a blob vs a mask
    rng = np.random.default_rng(5)
    pg = {}
    shapes = [(48, 64, 25), (96, 128, 25), (120, 97, 50), (64, 64, 10)]
    for i, (H, W, gs) in enumerate(shapes):
        yy, xx = np.mgrid[0:H, 0:W]
        blob = ((xx - W * 0.45) / (W * 0.3)) ** 2 + ((yy - H * 0.5) / (H * 0.35)) ** 2 <= 1.0
        ys, xs = np.nonzero(blob)
        sel = rng.choice(len(ys), size=min(len(ys), 40 + 30 * i), replace=False)
        pm = np.zeros((H, W), np.uint8)
        pm[ys[sel], xs[sel]] = 1
        mask = (np.roll(blob, (3, -4), axis=(0, 1))).astype(np.uint8)
        ext = mat.extend_pointgrid(torch.from_numpy(pm).bool(), gs).numpy()
        iou = mat.compute_point_mask_iou(torch.from_numpy(pm), torch.from_numpy(mask), gs)
        pg[f"pm{i}"], pg[f"mask{i}"], pg[f"grid{i}"], pg[f"ext{i}"], pg[f"iou{i}"] = pm, mask, gs, ext, np.float64(iou)
    # pred_tracks_to_binary_masks(return_mask=True) (cotracker_matching.py:453-503): hull fill, the 1-2 point fallback,
    # empty frames, NaN / out-of-bounds / half-integer coordinates
    H, W = 60, 90
    tr = rng.uniform(-6, 96, size=(2, 6, 40, 2)).astype(np.float32)
    tr[..., 1] = tr[..., 1] * (H / W)
    tr[0, 0, :10] = np.nan; tr[0, 1, 3] = (10.5, 20.5); tr[0, 1, 4] = (W - 0.5, 3.0)
    tr[1, 2] = -50.0                                   # nothing in bounds
    tr[1, 3, 2:] = 1e9; tr[1, 3, 0] = (5.2, 7.7); tr[1, 3, 1] = (80.0, 41.0)      # two points: discs
    tr[1, 4, 1:] = np.inf; tr[1, 4, 0] = (0.4, 0.4)                               # one point at the corner
    pg["hull_tracks"] = tr
    pg["hull_hw"] = np.asarray([H, W])
    pg["hull_masks"] = mat.pred_tracks_to_binary_masks(torch.from_numpy(tr), H, W, return_mask=True).numpy()
    pg["point_masks"] = mat.pred_tracks_to_binary_masks(torch.from_numpy(tr), H, W, return_mask=False).numpy()
    pg["n"] = len(shapes)
    pg["grid7"] = mat.get_points_on_a_grid(7, (48, 64)).numpy()
    path = os.path.join(ROOT, "tests", "golden", "pointgrid.npz")
    np.savez_compressed(path, **pg)
    print("wrote", path, [float(pg[f"iou{i}"]) for i in range(len(shapes))])


if __name__ == "__main__":
    main()
