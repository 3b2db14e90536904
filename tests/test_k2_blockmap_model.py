"""Host model of the votes kernel's sparse-tile mode (s2d_b200/csrc/point_votes.cu, `bm_done` branch of
point_votes_tab_kernel, and label_blockmap_kernel in label_vis.cu): block-summary maps (one byte per 4 x 4 pixels, 0xFF =
mixed, row pitch = ceil(W / 4) rounded up to a multiple of 4), the staged box of the map (rows start on a 32-bit word of
the map), the linear bitmap of the bounding box (bit = dy * bw + dx), the "first point on a pixel" claim and the label
lookup (block byte, exact label only for mixed blocks). The model follows the kernel statement by statement in unsigned
32-bit arithmetic and is compared with the oracle's sparse votes (oracle/keymask_oracle.py::point_votes =
cotracker_matching.py:453-503, 640-662)."""
import numpy as np
import pytest

from oracle import keymask_oracle as ko
from tests.test_k2_table_model import INVALID, pack

BUF_BYTES = 27904            # pv_buf_bytes(128, 8, 6, split=True)


def bm_pitch(W):
    return ((W + 3) // 4 + 3) & ~3


def bm_rows(H):
    return (H + 3) // 4


def blockmap_model(label):
    """label_blockmap_kernel: one frame -> u8 [bm_rows(H)][bm_pitch(W)]."""
    H, W = label.shape
    Hb, Wbp = bm_rows(H), bm_pitch(W)
    out = np.full((Hb, Wbp), 0xFF, np.uint8)
    for by in range(Hb):
        for bx in range((W + 3) // 4):
            blk = label[4 * by:4 * by + 4, 4 * bx:4 * bx + 4]          # ragged at the right / bottom edge
            first = blk[0, 0]
            if (blk == first).all() and first != 0xFF:
                out[by, bx] = first
    return out


def tile_votes_bm_model(tracks, label, bmap):
    """hits[256], uniq of one tile in sparse-tile mode; None when the kernel keeps the table path for the tile."""
    H, W = label.shape
    pk = pack(tracks, W, H)
    valid = pk != INVALID
    hist = np.zeros(256, np.int64)
    if not valid.any():
        return hist, 0
    ix, iy = (pk[valid] & 0xFFFF).astype(np.int64), (pk[valid] >> 16).astype(np.int64)
    x0, y0 = int(ix.min()), int(iy.min())
    bw, bh = int(ix.max()) + 1 - x0, int(iy.max()) + 1 - y0
    bx0, by0 = x0 >> 2, y0 >> 2
    bwb, bhb = ((x0 + bw - 1) >> 2) - bx0 + 1, ((y0 + bh - 1) >> 2) - by0 + 1
    a4 = bx0 & 3
    wpr = (a4 + bwb + 3) >> 2
    pitchb = 4 * wpr
    nblk = bhb * wpr
    nbw = (bh * bw + 31) >> 5
    blk_bytes = (4 * nblk + 15) & ~15
    if blk_bytes + 4 * nbw > BUF_BYTES:
        return None
    # staged copy: word i = (r, w) of the box, from map row by0 + r at byte bx0 - a4 + 4 w (inside the row: the pitch is a
    # multiple of 4 and covers ceil(W / 4))
    Wbp = bmap.shape[1]
    assert bx0 - a4 + pitchb <= Wbp and by0 + bhb <= bmap.shape[0]
    staged = bmap[by0:by0 + bhb, bx0 - a4:bx0 - a4 + pitchb].reshape(-1).copy()
    bits = np.zeros(nbw, np.uint32)
    pk0, lim = ((y0 << 16) + x0) & 0xFFFFFFFF, (bh << 16) & 0xFFFFFFFF
    flat = label.reshape(-1)
    blk_base = a4 - (by0 * pitchb + bx0)
    for p in pk:
        e = (int(p) - pk0) & 0xFFFFFFFF
        if e >= lim:
            assert int(p) == INVALID                                    # every valid point lies inside the box
            continue
        bit = (e >> 16) * bw + (e & 0xFFFF)
        assert (e & 0xFFFF) < bw and (bit >> 5) < nbw
        m = np.uint32(1 << (bit & 31))
        if bits[bit >> 5] & m:
            continue
        bits[bit >> 5] |= m
        x, y = int(p) & 0xFFFF, int(p) >> 16
        idx = blk_base + (y >> 2) * pitchb + (x >> 2)
        assert 0 <= idx < 4 * nblk
        lab = int(staged[idx])
        if lab == 0xFF:
            lab = int(flat[y * W + x])
        hist[lab] += 1
    return hist, int(hist.sum())


@pytest.mark.parametrize("H,W,seed", [(48, 64, 0), (33, 47, 1), (97, 131, 2), (30, 1021, 3), (5, 5, 4), (64, 66, 5)])
def test_blockmap_votes_model_vs_oracle(H, W, seed):
    rng = np.random.default_rng(seed)
    label = np.zeros((H, W), np.uint8)
    for k in range(1, 7):                                              # a few rectangles with ragged borders
        y, x = rng.integers(0, H), rng.integers(0, W)
        label[y:y + rng.integers(1, H // 2 + 2), x:x + rng.integers(1, W // 2 + 2)] = k
    label[rng.random((H, W)) < 0.02] = 7
    bmap = blockmap_model(label)
    assert bmap.shape == (bm_rows(H), bm_pitch(W)) and (bmap[:, (W + 3) // 4:] == 0xFF).all()
    # a uniform block's byte is the label of every pixel of the block
    for by in range(bm_rows(H)):
        for bx in range((W + 3) // 4):
            if bmap[by, bx] != 0xFF:
                assert (label[4 * by:4 * by + 4, 4 * bx:4 * bx + 4] == bmap[by, bx]).all()
    ran = 0
    for case in range(12):
        P = int(rng.integers(1, 300))
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        sx, sy = rng.uniform(1, W / 2), rng.uniform(1, H / 2)
        tr = np.stack([rng.normal(cx, sx, P), rng.normal(cy, sy, P)], axis=1).astype(np.float32)
        tr[::7] = tr[::-7][: len(tr[::7])]                             # duplicates
        if case % 3 == 0:
            tr[::5, 0] = np.nan
        if case % 4 == 1:
            tr[:, :] = np.round(tr) + 0.5                              # half-integers round to even
        if case == 11:
            tr[:, 0] = -50                                             # nothing inside the frame
        got = tile_votes_bm_model(tr, label, bmap)
        if got is None:
            continue
        ran += 1
        h, u = ko.point_votes(tr[None], label[None], 0, 0, nbins=256)
        assert got[1] == int(u[0]) and np.array_equal(got[0], h[0]), case
    assert ran >= 6


def test_blockmap_box_too_large_keeps_the_table_path():
    H, W = 1080, 1920
    label = np.zeros((H, W), np.uint8)
    tr = np.asarray([[0, 0], [W - 1, H - 1]], np.float32)
    assert tile_votes_bm_model(tr, label, np.zeros((bm_rows(H), bm_pitch(W)), np.uint8)) is None
