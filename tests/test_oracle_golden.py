"""The CPU restatement (oracle/keymask_oracle.py) against fixtures produced by the unmodified
reference (oracle/make_golden.py). Runs without a GPU and without /root/reference."""
import numpy as np
import pytest

from oracle import keymask_oracle as ko
from oracle.compare import check_against_golden


def test_oracle_matches_reference_golden(golden_case):
    name, g, labels, tracks, vis = golden_case
    res = ko.discover(labels, tracks, vis, g["visibility_threshold"], g["matching_threshold"])
    check_against_golden(res, g)


@pytest.mark.parametrize("name", ["long", "long_thr"])
def test_goldens_longer_than_one_bit_word_end_with_status_1(name):
    """36- and 40-frame videos that end with status 1 in the unmodified reference: more than one 32-bit word per
    visibility / match bit row. They sit in tests/golden/ so that the -m gpu tests run select / group with TW = 2."""
    from tests.conftest import load_golden
    g, labels, tracks, vis = load_golden(name)
    assert g["status"] == 1 and labels.shape[0] > 32


def test_candidate_files_match(golden_case):
    name, g, labels, tracks, vis = golden_case
    res = ko.discover(labels, tracks, vis, g["visibility_threshold"], g["matching_threshold"])
    files = sorted(f"{folder}/cluster{idx}_frame{f}_mask{m}.png"
                   for folder, idx, cands in ko.lexsorted_cluster_folders(res["clusters"]) for f, m in cands)
    assert files == g["candidate_files"]
    if g["status"] == 1:
        gf = sorted(f"cluster_{gr['cluster_id']}/group_{lab}/frame{f}_mask{m}.png"
                    for gr in res["groupings"] for lab, fms in gr["overall_mask_ids_per_label"].items()
                    for f, m in fms)
        assert gf == g["group_files"]


def test_tracker_requests_match(golden_case):
    """grid size heuristic and backward_tracking flag of the stage-D tracker request
    (cotracker_matching.py:1064-1073)."""
    name, g, labels, tracks, vis = golden_case
    if g["status"] != 1:
        pytest.skip("video fails in the reference")
    res = ko.discover(labels, tracks, vis, g["visibility_threshold"], g["matching_threshold"])
    nm = len(res["query_frame"])
    stage_d = g["tracker_calls"][nm:]
    got = [[q["frame_id"], q["mask_id"], q["grid_size"], bool(q["backward_tracking"])] for q in res["queries"]]
    assert got == [[c[0], c[1], c[2], bool(c[3])] for c in stage_d]


@pytest.mark.parametrize("eps,ms", [(0.2, 5), (0.1, 5), (0.1, 3), (0.05, 5)])
def test_dbscan_restatement_vs_sklearn(eps, ms):
    from sklearn.cluster import DBSCAN
    rng = np.random.default_rng(7)
    for it in range(60):
        n = int(rng.integers(1, 60))
        d = int(rng.integers(1, 70))
        base = rng.random((int(rng.integers(1, 5)), d)) < 0.5
        X = base[rng.integers(0, len(base), n)] ^ (rng.random((n, d)) < rng.choice([0.0, 0.03, 0.1]))
        if it % 3 == 0:
            X[rng.integers(0, n, max(1, n // 4))] = False
        for arr in (X, X.astype(np.float32)):
            ref = DBSCAN(eps=eps, min_samples=ms, metric="hamming").fit(arr).labels_
            assert np.array_equal(ref, ko.dbscan_hamming(arr, eps, ms)), (it, n, d)


def test_dense_port_matches_reference_pairs(golden_case):
    """the dense torch-CPU port used as bench.py's CPU baseline reproduces the reference's
    per-pair (intersection, union, iou) and match lists."""
    import torch
    from oracle import dense_port
    name, g, labels, tracks, vis = golden_case
    if g["status"] != 1:
        pytest.skip("video fails in the reference")
    T, H, W = labels.shape
    lab = torch.from_numpy(labels.astype(np.int64))[..., None]
    qs = g["queries"][:: max(1, len(g["queries"]) // 4)]
    for q in qs:
        row = q["overall_mask_id"]
        m, comps, o2 = dense_port.match_query_dense(lab, torch.from_numpy(tracks[row]), q["v_range"][0],
                                                    q["v_range"][1], H, W, g["matching_threshold"])
        assert [[c[0], c[1], c[2], c[3]] for c in comps] == [[c[0], c[1], c[3], c[4]] for c in q["comps"]]
        assert o2 == q["one2x"]
        gid = {(c[0], c[1]): c[2] for c in q["comps"]}
        assert [gid[fm] for fm in m] == q["matches"]
    V = dense_port.visibility_rows(torch.from_numpy(vis.astype(bool))).numpy()
    ref = np.asarray([r["visibility"] for r in g["visibility_rows"]], np.float32)
    assert np.array_equal(V, ref)


def test_appearance_events_oracle_vs_reference_golden():
    """numpy restatement of extract_appearance_events against vectors produced by the unmodified
    reference function (oracle/make_golden_events.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "events.npz"))
    for name in g["names"]:
        V = g[f"{name}__V"]
        sw, th, k = g[f"{name}__params"]
        ev, _ = ko.appearance_events(V, int(sw), float(th), int(k))
        npairs, pairs = g[f"{name}__npairs"], g[f"{name}__pairs"]
        assert [len(ev[i]) for i in range(V.shape[0])] == npairs.tolist(), name
        flat = [p for i in range(V.shape[0]) for p in ev[i]]
        assert np.array_equal(np.asarray(flat, np.int32).reshape(-1, 2), pairs), name


def test_coco_rle_restatement_known_answers_and_roundtrip():
    """oracle/coco_rle.py against hand-derived COCO answers (column-major runs, leading zero run) and the
    product's string codec (rleToString / rleFrString) as a round trip."""
    from oracle import coco_rle as cr
    from s2d_b200.keymask_ident.annotations import coco_rle as codec
    m = np.array([[0, 1, 1], [0, 1, 0]], np.uint8)             # columns: 00 | 11 | 10
    assert cr.counts(m) == [2, 3, 1] and cr.area(m) == 3 and cr.bbox(m) == [1.0, 0.0, 2.0, 2.0]
    m1 = np.array([[1, 0], [1, 1]], np.uint8)                  # starts set: leading zero-length run
    assert cr.counts(m1) == [0, 2, 1, 1]
    assert cr.counts(np.zeros((3, 4), np.uint8)) == [12] and cr.bbox(np.zeros((3, 4))) == [0.0] * 4
    rng = np.random.default_rng(3)
    for h, w in [(1, 1), (5, 7), (33, 64), (120, 97)]:
        mm = (rng.random((h, w)) < 0.4).astype(np.uint8)
        c = cr.counts(mm)
        assert sum(c) == h * w and np.array_equal(cr.decode(c, h, w), mm)
        assert codec.from_string(codec.to_string(c)) == c
    # hand-derived strings (maskApi.c rleToString: 5 data bits per character from the low end, bit 0x20 = "more",
    # offset 48; counts after the third are stored as the difference to the count two places back):
    #   2, 3, 1 -> '2' '3' '1';   0, 2, 1, 1 -> last is 1 - 2 = -1 = 0b11111 with the sign bit set, no more -> chr(31 + 48) = 'O'
    #   5, 4, 3, 2 -> last is 2 - 4 = -2 -> chr(30 + 48) = 'N';   100 = 0b00011_00100 -> chr(4 + 32 + 48) = 'T', chr(3 + 48) = '3'
    for cnts, text in [([2, 3, 1], "231"), ([0, 2, 1, 1], "021O"), ([5, 4, 3, 2], "543N"), ([100], "T3"),
                       ([12], "<"), ([0, 16], "0`0")]:
        assert codec.to_string(cnts) == text and codec.from_string(text) == cnts
    big = [0, 5, 100000, 3, 70000, 1, 1, 40]                   # multi-character counts and negative deltas
    assert codec.from_string(codec.to_string(big)) == big


def test_coco_rle_counts_vs_third_party_encoder():
    """The run counts of oracle/coco_rle.py (the comparator of the GPU encoder, row f2) against an implementation that is
    not ours: transformers' SAM post-processing `_mask_to_rle` ("in the format expected by pycoco tools": column-major
    runs, leading zero run, `size` = [h, w]) - the uncompressed form pycocotools' frPyObjects accepts. pycocotools itself
    is not installable here, so this pins the counts to a second, independent restatement of the COCO convention; the
    string codec stays pinned by the hand-derived answers above."""
    torch = pytest.importorskip("torch")
    sam = pytest.importorskip("transformers.models.sam.image_processing_sam")
    from oracle import coco_rle as cr
    rng = np.random.default_rng(11)
    masks = [np.zeros((7, 5), np.uint8), np.ones((7, 5), np.uint8)]
    for h, w, p in [(1, 1, 0.5), (5, 7, 0.4), (33, 64, 0.1), (120, 97, 0.7), (64, 64, 0.02)]:
        masks.append((rng.random((h, w)) < p).astype(np.uint8))
    blob = np.zeros((90, 130), np.uint8)
    yy, xx = np.mgrid[:90, :130]
    blob[((yy - 40) / 25.0) ** 2 + ((xx - 70) / 40.0) ** 2 <= 1.0] = 1      # an object-shaped mask
    blob[0, 0] = 1                                                           # starts set: leading zero-length run
    masks.append(blob)
    for m in masks:
        theirs = sam._mask_to_rle(torch.from_numpy(m.astype(bool))[None])[0]
        assert theirs["size"] == list(m.shape)
        assert [int(c) for c in theirs["counts"]] == cr.counts(m), m.shape
        assert cr.area(m) == int(m.sum()) == sum(int(c) for c in theirs["counts"][1::2])


def test_colour_to_label_rule_vs_reference_golden():
    """f1: the product's host rule for colour-coded mask frames (rank of the RGB tuple among the frame's non-black
    colours, lexicographic) against outputs of the unmodified reference's convert_lblimg_to_maskid
    (crw_utils.py:688-711; fixture from oracle/make_golden_colors.py). The GPU kernel is compared with this host rule
    in tests/test_gpu_parity.py::test_color_to_labels_vs_host_rule."""
    import os
    from s2d_b200.keymask_ident.crw_utils import convert_lblimg_to_maskid, rgb_to_label_ids
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_cpu", "colors.npz"))
    n = len([k for k in z.files if k.startswith("rgb")])
    assert n >= 5
    for i in range(n):
        rgb, want = z[f"rgb{i}"], z[f"ids{i}"]
        assert np.array_equal(rgb_to_label_ids(rgb), want[..., 0])
        got = np.asarray(convert_lblimg_to_maskid(rgb))
        assert got.shape == want.shape and np.array_equal(got, want)
