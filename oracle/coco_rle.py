"""TEST INFRASTRUCTURE. numpy restatement of the COCO mask API routines the reference's stage E calls
through pycocotools (annotations.py:100-106, convert_results_to_annotations.py:80-81): pycocotools
(un-pinned, not installable here) -> cocodataset/cocoapi common/maskApi.c rleEncode, rleArea,
rleToBbox, rleDecode. Parity unpinned against the library itself; pinned against its published
algorithm and by round trips."""
from __future__ import annotations

import numpy as np


def counts(mask: np.ndarray):
    """rleEncode: column-major run lengths starting with a run of zeros (possibly empty)."""
    flat = np.asarray(mask, dtype=np.uint8).reshape(-1, order="F") != 0
    if flat.size == 0:
        return [0]
    change = np.nonzero(flat[1:] != flat[:-1])[0] + 1
    edges = np.concatenate(([0], change, [flat.size]))
    runs = np.diff(edges).tolist()
    if flat[0]:
        runs = [0] + runs
    return runs


def area(mask: np.ndarray) -> int:
    """rleArea: sum of the odd runs."""
    return int(sum(counts(mask)[1::2]))


def bbox(mask: np.ndarray):
    """rleToBbox: (x, y, w, h) of the set pixels, zeros for an empty mask."""
    m = np.asarray(mask) != 0
    if not m.any():
        return [0.0, 0.0, 0.0, 0.0]
    ys, xs = np.nonzero(m.any(axis=1))[0], np.nonzero(m.any(axis=0))[0]
    return [float(xs[0]), float(ys[0]), float(xs[-1] - xs[0] + 1), float(ys[-1] - ys[0] + 1)]


def decode(cnts, h: int, w: int) -> np.ndarray:
    """rleDecode."""
    flat = np.zeros(h * w, np.uint8)
    p, v = 0, 0
    for c in cnts:
        flat[p:p + c] = v
        p += c
        v ^= 1
    return flat.reshape((h, w), order="F")


def rle_area(cnts) -> int:
    """rleArea on the counts."""
    return int(sum(cnts[1::2]))


def rle_to_bbox(cnts, h: int):
    """rleToBbox, loop for loop (maskApi.c): boundaries of the first 2*floor(n/2) runs."""
    m = (len(cnts) // 2) * 2
    if m == 0:
        return [0.0, 0.0, 0.0, 0.0]
    xs = ys = None
    xe = ye = 0
    cc, xp = 0, 0
    full = False
    for j in range(m):
        cc += int(cnts[j])
        t = cc - (j % 2)
        y = t % h
        x = (t - y) // h
        if j % 2 == 0:
            xp = x
        elif xp < x:
            full = True
        xs = x if xs is None else min(xs, x)
        ys = y if ys is None else min(ys, y)
        xe, ye = max(xe, x), max(ye, y)
    if full:
        ys, ye = 0, h - 1
    return [float(xs), float(ys), float(xe - xs + 1), float(ye - ys + 1)]
