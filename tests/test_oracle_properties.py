"""Properties of the oracle's building blocks, checked on the host against torch CPU (the library the reference
computes with) on seeded random inputs: the rules SURVEY.md Appendix A lists as deciding bit-exactness."""
import re

import numpy as np
import pytest
import torch

from oracle import keymask_oracle as ko


def test_round_tracks_is_torch_round_long_with_specials():
    """pred_tracks.round().long() (cotracker_matching.py:472): half-to-even in float32, NaN / inf out of every frame."""
    rng = np.random.default_rng(11)
    xy = (rng.random((4000, 2), dtype=np.float32) * 900 - 20).astype(np.float32)
    xy[:600] = np.floor(xy[:600]) + np.float32(0.5)              # exact .5 boundaries, both parities
    xy[600:620] = [[-0.5, 0.5], [853.5, 479.5]] * 10
    want = torch.from_numpy(xy).round().long().numpy()
    assert np.array_equal(ko.round_tracks(xy), want)
    bad = np.array([[np.nan, 1.0], [np.inf, 2.0], [-np.inf, 3.0], [1e30, 4.0]], np.float32)
    r = ko.round_tracks(bad)
    assert (r[:, 0] < 0).all() and np.array_equal(r[:, 1], [1, 2, 3, 4])     # dropped by the 0 <= x test


@pytest.mark.parametrize("seed,H,W,P,M", [(1, 40, 56, 300, 5), (2, 33, 47, 900, 12), (3, 64, 64, 5000, 3)])
def test_sparse_votes_equal_dense_overlap(seed, H, W, P, M):
    """hits / uniq of the sparse restatement == intersection / union of the reference's dense formulation
    (pred_tracks_to_binary_masks + compute_point_mask_intersection, cotracker_matching.py:453-503, 640-662),
    written here with torch ops on full frames as the reference does."""
    rng = np.random.default_rng(seed)
    T = 4
    labels = rng.integers(0, M + 1, size=(T, H, W)).astype(np.uint8)
    labels[:, : H // 4] = 0
    tracks = (rng.random((T, P, 2), dtype=np.float32) * np.float32([W + 6, H + 6]) - 3).astype(np.float32)
    tracks[:, : P // 5] = np.floor(tracks[:, : P // 5]) + np.float32(0.5)
    tracks[:, P // 5: P // 4] = tracks[:, :1]                                 # duplicates collapse
    hits, uniq = ko.point_votes(tracks, labels, 0, T - 1)
    for t in range(T):
        pts = torch.from_numpy(tracks[t]).round().long()
        ok = (pts[:, 0] >= 0) & (pts[:, 0] < W) & (pts[:, 1] >= 0) & (pts[:, 1] < H)
        pm = torch.zeros((H, W), dtype=torch.uint8)
        pm[pts[ok, 1], pts[ok, 0]] = 1
        assert np.array_equal(ko.rasterise_tracks(tracks[t], H, W), pm.numpy())
        for m in range(M + 1):
            mask = torch.from_numpy((labels[t] == m).astype(np.uint8))
            mp = mask & pm                                                     # mask restricted to the points
            inter = int((pm & mp).sum())
            union = int((pm | mp).sum())
            assert hits[t, m] == inter and uniq[t] == union
            assert ko.iou_of(inter, union) == (inter / union if union else 0.0)


def test_visibility_mean_is_torch_float32_mean():
    """torch.mean(pred_visibility.float(), dim=2) (cotracker_occlusions.py:359) for point counts that are not
    powers of two: float32(count) / float32(P), not count * (1 / P)."""
    rng = np.random.default_rng(5)
    for P in (1, 3, 7, 100, 1000, 4096, 4099):
        vis = rng.random((6, 9, P)) < rng.random((6, 9, 1))
        want = torch.mean(torch.from_numpy(vis).float(), dim=2).numpy()
        got = ko.visibility_mean(vis)
        assert got.dtype == np.float32 and np.array_equal(got, want)


def test_hamming_kmax_is_the_float64_threshold():
    for D in list(range(1, 80)) + [300, 720, 1024]:
        for eps in (0.05, 0.1, 0.2, 0.3):
            k = ko.hamming_kmax(D, eps)
            assert k >= 0 and float(k) / float(D) <= eps
            assert k == D or float(k + 1) / float(D) > eps


def test_visible_ranges_are_the_maximal_runs_of_ones():
    rng = np.random.default_rng(9)
    for n in (1, 2, 31, 32, 33, 64, 100):
        for _ in range(40):
            v = rng.random(n) < rng.random()
            s = "".join("1" if b else "0" for b in v)
            want = [(m.start(), m.end() - 1) for m in re.finditer("1+", s)]
            assert ko.visible_ranges(v) == want


def test_object_enumeration_drops_the_smallest_label_even_without_background():
    """unique(...)[1:] (cotracker_matching.py:294): the smallest present label is dropped whether or not it is 0."""
    labels = np.array([[[0, 2, 2], [5, 5, 0]], [[3, 3, 4], [4, 7, 7]]], np.uint8)
    ids = ko.enumerate_objects(labels)
    assert [list(x) for x in ids] == [[2, 5], [4, 7]]
    qf, ql, lut = ko.global_id_lookup(labels)
    assert list(qf) == [0, 0, 1, 1] and list(ql) == [2, 5, 4, 7] and lut[(1, 7)] == 3
    for t in range(2):
        assert np.array_equal(ids[t], torch.unique(torch.from_numpy(labels[t].astype(np.int64)))[1:].numpy())


def test_pack_vis_bits_layout():
    """wire format of S2D_DESC_VIS_BITS: bit p % 32 of int32 word p // 32 is flag p (little-endian bit order)."""
    import torch
    from s2d_b200.pipeline import pack_vis_bits
    rng = np.random.default_rng(3)
    for P in (1, 31, 32, 33, 100, 4096):
        vis = (rng.random((3, 4, P)) < 0.5).astype(np.uint8)
        w = pack_vis_bits(torch.from_numpy(vis)).numpy()
        assert w.dtype == np.int32 and w.shape == (3, 4, (P + 31) // 32)
        ref = np.packbits(np.pad(vis, ((0, 0), (0, 0), (0, (-P) % 32))), axis=2, bitorder="little")
        assert np.array_equal(w.view(np.uint8).reshape(3, 4, -1), ref)
