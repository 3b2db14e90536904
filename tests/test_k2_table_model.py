"""Host model of the label-table votes kernel's integer arithmetic (s2d_b200/csrc/point_votes.cu,
point_votes_tab_kernel): packed coordinates, bounding box, table pitch for both fetch modes (2D TMA boxes / one bulk
copy per row with the pitch chosen = W mod 16), bands, the wrap-around band test `e < lim`, the offset
`(e >> 16) * (pitch - 65536) + e`, and claim-by-OR with 0xFF as "already taken". The model follows the kernel
statement by statement in uint32 arithmetic and is compared with the oracle's sparse votes
(oracle/keymask_oracle.py::point_votes = cotracker_matching.py:453-503, 640-662). It documents why the row copies of
the unaligned mode never overwrite each other's payload, whatever order they land in."""
import numpy as np
import pytest

from oracle import keymask_oracle as ko

U32 = 0xFFFFFFFF
BUF_BYTES = 35712            # pv_buf_bytes(128, 32, 6)
MAX_BANDS = 6                # PV_MAX_BANDS
TMAPS = 32                   # S2D_PV_TMAPS: box widths 16 .. 512 pixels
INVALID = U32


def pack(tracks, W, H):
    """pv_pack: __float2int_rn(fmaxf(v, -1)) as uint32, (iy << 16) + ix when inside the frame."""
    v = tracks.astype(np.float32)
    v = np.where(np.isnan(v), np.float32(-1), np.maximum(v, np.float32(-1)))      # fmaxf(NaN, -1) = -1
    r = np.clip(np.rint(v.astype(np.float64)), -2**31, 2**31 - 1).astype(np.int64)  # F2I saturates; rint = half-even
    ix, iy = r[:, 0] & U32, r[:, 1] & U32
    ok = (ix < W) & (iy < H)
    return np.where(ok, ((iy << 16) + ix) & U32, INVALID).astype(np.uint64)


def tile_votes_model(tracks, label, *, tma, base_mod16=0, rng=None):
    """hits[256], uniq of one (query, frame) tile as the kernel computes them in table mode; None when the kernel
    would take its in-kernel fallback (bounding box taller than MAX_BANDS bands)."""
    H, W = label.shape
    flat = label.reshape(-1)
    pk = pack(tracks, W, H)
    valid = pk != INVALID
    hist = np.zeros(256, np.int64)
    if not valid.any():
        return hist, 0
    ix, iy = (pk[valid] & 0xFFFF).astype(np.int64), (pk[valid] >> 16).astype(np.int64)
    x0, y0 = int(ix.min()), int(iy.min())
    bw, bh = int(ix.max()) + 1 - x0, int(iy.max()) + 1 - y0
    use_tma = tma and W % 16 == 0 and base_mod16 == 0 and bw + (x0 & 15) <= 16 * TMAPS
    if use_tma:
        pitch = (bw + (x0 & 15) + 15) & ~15
        R = (BUF_BYTES // pitch) & ~15
    else:
        pitch = bw + 15
        pitch += (W - pitch) & 15                       # pitch = W (mod 16), pitch >= bw + 15
        R = BUF_BYTES // pitch
    if R * MAX_BANDS < bh:
        return None
    for b0 in range(0, bh, R):
        rows = min(R, bh - b0)
        tab = np.full(BUF_BYTES + 64, 0xEE, np.int64)   # 0xEE: bytes no copy wrote (must never be claimed)
        if use_tma:
            a15 = x0 & 15
            nbox = (rows + 15) >> 4
            assert nbox * 16 * pitch <= BUF_BYTES and pitch % 16 == 0
            c0 = x0 & ~15
            for r in range(nbox * 16):                  # boxes of 16 rows x pitch pixels; out-of-range -> zero fill
                y = y0 + b0 + r
                row = np.zeros(pitch, np.int64)
                if y < H:
                    n = max(0, min(pitch, W - c0))
                    row[:n] = flat[y * W + c0: y * W + c0 + n]
                tab[r * pitch: (r + 1) * pitch] = row
        else:
            A = base_mod16 + (y0 + b0) * W + x0         # address of the band's first pixel (mod 16 is all that matters)
            a15 = A & 15
            order = list(range(rows))
            if rng is not None:
                rng.shuffle(order)                      # bulk copies complete in any order
            payload = np.zeros(BUF_BYTES + 64, bool)
            for r in order:
                g = A + r * W
                ph = g & 15
                d0, ln = r * pitch + a15 - ph, (ph + bw + 15) & ~15
                assert d0 % 16 == 0 and d0 >= 0 and d0 + ln <= BUF_BYTES + 64 and (g - ph) % 16 == 0
                src = (y0 + b0 + r) * W + x0 - ph       # pixel index of the first copied byte (may lie left of the box)
                idx = np.arange(src, src + ln)
                inside = (idx >= 0) & (idx < flat.size)  # the real copy reads neighbouring bytes of the allocation
                seg = np.where(inside, flat[np.clip(idx, 0, flat.size - 1)], 0)
                assert not payload[d0:d0 + ln].any(), "a row copy would overwrite another row's payload"
                tab[d0:d0 + ln] = seg
                payload[r * pitch + a15: r * pitch + a15 + bw] = True
        pk0 = (((y0 + b0) << 16) + x0) & U32
        lim = rows << 16
        negc = (pitch - 65536) & U32
        for p in pk:                                    # every point of the tile, in order; invalid ones fail e < lim
            e = (int(p) - pk0) & U32
            if e >= lim:
                hist[255] += 1                          # the kernel's dummy word: reads back 0xFF
                continue
            off = (((e >> 16) * negc) + e + a15) & U32
            assert off < rows * pitch + a15 and off < BUF_BYTES
            lab = int(tab[off])
            assert lab != 0xEE, "claimed a byte no copy wrote"
            hist[lab] += 1
            tab[off] = 0xFF                             # atomicOr(word, 0xFF << 8 * (off & 3))
    hits = hist.copy()
    hits[255] = 0
    return hits, int(hits.sum())


@pytest.mark.parametrize("H,W,P,spread", [(720, 1280, 4096, (260, 170)), (480, 854, 4096, (200, 150)), (97, 333, 777, (90, 60)),
                                          (1080, 1920, 4096, (600, 420)), (720, 1280, 4096, (400, 300)), (64, 64, 2000, (64, 64)), (300, 1001, 500, (30, 280))])
@pytest.mark.parametrize("tma", [False, True])
def test_table_model_equals_oracle_votes(H, W, P, spread, tma):
    rng = np.random.default_rng(H * 7 + W + P + int(tma))
    label = rng.integers(0, 21, size=(H, W)).astype(np.uint8)
    for trial in range(4):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        t = np.stack([rng.uniform(cx - spread[0] / 2, cx + spread[0] / 2, P),
                      rng.uniform(cy - spread[1] / 2, cy + spread[1] / 2, P)], axis=1).astype(np.float32)
        t[: P // 10] = np.floor(t[: P // 10]) + np.float32(0.5)             # half-integer coordinates
        t[P // 10: P // 8] = t[:1]                                           # duplicates
        t[-4:] = [[np.nan, 3.0], [np.inf, 5.0], [-7.5, 2.0], [3.0, 1e9]]     # never inside the frame
        out = tile_votes_model(t, label, tma=tma, base_mod16=int(rng.integers(0, 16)) if not tma else 0, rng=rng)
        if out is None:
            continue
        hits, uniq = out
        rh, ru = ko.point_votes(t[None], label[None], 0, 0)
        assert np.array_equal(hits, rh[0]) and uniq == ru[0]


def test_table_model_handles_empty_and_single_pixel_tiles():
    label = np.arange(40 * 48, dtype=np.int64).reshape(40, 48) % 7
    label = label.astype(np.uint8)
    none_inside = np.array([[-3.0, 2.0], [100.0, 3.0], [np.nan, np.nan]], np.float32)
    hits, uniq = tile_votes_model(none_inside, label, tma=False)
    assert uniq == 0 and hits.sum() == 0
    one = np.array([[47.4, 39.4]] * 5, np.float32)                           # five tracks on the last pixel
    for tma in (False, True):
        hits, uniq = tile_votes_model(one, label, tma=tma)
        assert uniq == 1 and hits[label[39, 47]] == 1
