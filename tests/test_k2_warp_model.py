"""Host model of the one-warp-per-tile votes kernel's integer arithmetic (s2d_b200/csrc/point_votes.cu,
point_votes_warp_kernel: tiles of <= 1024 points): packed coordinates, bounding box from packed 16-bit min / max, the
bitmap of the box in bands of PVW_BITS pixels (bit = dy * bw + dx - base in wrapping unsigned arithmetic, `e < lim` as the
"inside the frame" test), claim of the first point per pixel, label address org + dy * W + dx, ragged last row of points
(n < P). The model follows the kernel statement by statement in uint32 arithmetic and is compared with the oracle's sparse
votes (oracle/keymask_oracle.py::point_votes = cotracker_matching.py:453-503, 640-662)."""
import numpy as np
import pytest

from oracle import keymask_oracle as ko
from tests.test_k2_table_model import INVALID, pack

U32 = 0xFFFFFFFF


def tile_votes_warp_model(tracks, label, n=None, bits_per_band=32768):
    """hits[256], uniq, number of bands of one (query, frame) tile as point_votes_warp_kernel computes them."""
    H, W = label.shape
    P = len(tracks)
    n = P if n is None else n
    assert P <= 1024 and P % 2 == 0
    flat = label.reshape(-1)
    pk_all = pack(tracks, W, H).astype(np.uint64)
    # lane l holds points 64 k + 2 l, + 1; rows k < kfull are complete, row kfull is ragged, rows >= kmax are invalid
    kfull, kmax = n >> 6, (n + 63) >> 6
    pk = np.full(1024, INVALID, np.uint64)
    for k in range(16):
        for lane in range(32):
            p0 = 2 * (k * 32 + lane)
            if k < kfull:
                pk[p0], pk[p0 + 1] = pk_all[p0], pk_all[p0 + 1]
            elif k < kmax:
                if p0 < n:
                    pk[p0] = pk_all[p0]
                if p0 + 1 < n:
                    pk[p0 + 1] = pk_all[p0 + 1]
    hist = np.zeros(256, np.int64)
    # packed min / max per 16-bit half; an invalid point is 0xFFFF for the min and wraps to (1, 0) for the max
    lo, hi = pk & 0xFFFF, pk >> 16
    x0, y0 = int(lo.min()), int(hi.min())
    pmx = (pk + 0x00010001) & U32
    x1, y1 = int((pmx & 0xFFFF).max()), int((pmx >> 16).max())
    if ((y0 << 16) | x0) == U32:
        return hist, 0, 0
    bw, bh = x1 - x0, y1 - y0
    assert bw >= 1 and bh >= 1 and x0 + bw <= W and y0 + bh <= H
    pk0, lim = ((y0 << 16) + x0) & U32, (bh << 16) & U32
    npx, org = bw * bh, y0 * W + x0
    bands = 0
    for base in range(0, npx, bits_per_band):
        bands += 1
        bits = np.zeros(bits_per_band // 32, np.uint32)
        for g in range(0, 32, 8):                       # the kernel's group order: rows g / 2 .. g / 2 + 3 of every lane
            if g // 2 >= kmax:
                break
            for lane in range(32):
                for k in range(g, g + 8):
                    p = int(pk[2 * ((k // 2) * 32 + lane) + (k & 1)])
                    e = (p - pk0) & U32
                    lin = ((e >> 16) * bw + (e & 0xFFFF) - base) & U32
                    if not (e < lim and lin < bits_per_band):
                        continue
                    assert (e & 0xFFFF) < bw
                    m = np.uint32(1 << (lin & 31))
                    if bits[lin >> 5] & m:
                        continue
                    bits[lin >> 5] |= m
                    hist[int(flat[org + (e >> 16) * W + (e & 0xFFFF)])] += 1
    return hist, int(hist.sum()), bands


@pytest.mark.parametrize("H,W,seed", [(48, 64, 0), (33, 47, 1), (97, 131, 2), (30, 1021, 3), (5, 5, 4), (300, 400, 5)])
def test_warp_votes_model_vs_oracle(H, W, seed):
    rng = np.random.default_rng(seed)
    label = rng.integers(0, 256, size=(H, W)).astype(np.uint8)         # every label id, 255 included
    label[: H // 2, : W // 2] = 3
    multi = 0
    for case in range(10):
        P = int(rng.integers(1, 200)) * 2
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        sx, sy = rng.uniform(1, W / 2), rng.uniform(1, H / 2)
        tr = np.stack([rng.normal(cx, sx, P), rng.normal(cy, sy, P)], axis=1).astype(np.float32)
        tr[::7] = tr[::-7][: len(tr[::7])]                             # duplicates
        if case % 3 == 0:
            tr[::5, 0] = np.nan
        if case % 4 == 1:
            tr[:, :] = np.round(tr) + 0.5                              # half-integers round to even
        if case == 9:
            tr[:, 0] = -50                                             # nothing inside the frame
        n = P if case % 2 else int(rng.integers(0, P + 1))             # ragged: only the first n points count
        for bpb in (32768, 256):                                       # the kernel's band size and a small one (many bands)
            h, u, bands = tile_votes_warp_model(tr, label, n=n, bits_per_band=bpb)
            multi += bands > 1
            ho, uo = ko.point_votes(tr[None, :n], label[None], 0, 0, nbins=256)
            assert u == int(uo[0]) and np.array_equal(h, ho[0]), (case, bpb)
    assert multi >= 5 or H * W <= 256


def test_warp_votes_model_full_tile_and_large_box():
    """1024 points spread over a 1080p frame: 64 bands of 32768 pixels."""
    rng = np.random.default_rng(7)
    H, W = 1080, 1920
    label = rng.integers(0, 30, size=(H, W)).astype(np.uint8)
    tr = np.stack([rng.uniform(-5, W + 5, 1024), rng.uniform(-5, H + 5, 1024)], axis=1).astype(np.float32)
    h, u, bands = tile_votes_warp_model(tr, label)
    assert bands > 50
    ho, uo = ko.point_votes(tr[None], label[None], 0, 0, nbins=256)
    assert u == int(uo[0]) and np.array_equal(h, ho[0])
