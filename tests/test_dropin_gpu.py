"""The file-based drop-in modules (s2d_b200/keymask_ident) driven in the reference driver's order on
an on-disk video, compared with what the unmodified reference wrote for the same video (goldens)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

# (every test but the host-only hull raster test is marked gpu individually below)


def _write_video(root, labels, name="vid0"):
    import cv2
    T, H, W = labels.shape
    vdir = os.path.join(root, "ytvis2021", "train", "JPEGImages", name)
    mdir = os.path.join(root, "masks", name)
    os.makedirs(vdir); os.makedirs(mdir)
    pal = np.zeros((256, 3), np.uint8)
    for l in range(1, 256):      # strictly increasing lexicographically: rank order == label order
        pal[l] = (10 + l // 2, (37 * l + 11) % 256, (91 * l + 5) % 256) if l > 1 else (10, 1, 1)
    pal[1:, 0] = 10 + np.arange(1, 256) // 2
    pal[1:, 1] = (np.arange(1, 256) % 2) * 100 + 7
    for t in range(T):
        rgb = pal[labels[t]]
        cv2.imwrite(os.path.join(mdir, f"{t:05d}.png"), rgb[..., ::-1])
        cv2.imwrite(os.path.join(vdir, f"{t:05d}.jpg"), np.full((H, W, 3), 127, np.uint8))
    return vdir, mdir


class _Replay:
    """CoTrackerPredictor stand-in replaying the fixture's tracks / visibility."""

    def __init__(self, labels, tracks, vis):
        from oracle.keymask_oracle import global_id_lookup
        self.labels, self.tracks, self.vis = labels, tracks, vis
        _, _, self.lut = global_id_lookup(labels)
        self.calls = []

    def __call__(self, checkpoint=None):
        return self

    def cuda(self):
        return self

    def predict(self, video, grid_size, grid_query_frame, segm_mask, backward_tracking):
        sm = segm_mask[0, 0].cpu().numpy()
        ys, xs = np.nonzero(sm)
        lab = int(self.labels[grid_query_frame][ys[0], xs[0]])
        q = self.lut[(int(grid_query_frame), lab)]
        self.calls.append([int(grid_query_frame), lab, int(grid_size), bool(backward_tracking)])
        return torch.from_numpy(self.tracks[q][None].copy()), torch.from_numpy(self.vis[q][None].astype(bool))


class _Model:
    def __init__(self, rp):
        self.rp = rp

    def cuda(self):
        return self

    def __call__(self, video, grid_size=0, grid_query_frame=0, segm_mask=None, backward_tracking=False):
        return self.rp.predict(video, grid_size, grid_query_frame, segm_mask, backward_tracking)


@pytest.mark.gpu
def test_dropin_stages_match_reference_files(golden_case, tmp_path):
    from oracle.compare import _close
    from s2d_b200.keymask_ident import (_engine, cotracker_matching, cotracker_occlusions, crw_utils,
                                       identify_visibility_windows, keymask_utils)
    name, g, labels, tracks, vis = golden_case
    root = str(tmp_path)
    vdir, mdir = _write_video(root, labels)
    rp = _Replay(labels, tracks, vis)
    _engine.set_tracker_factory(lambda checkpoint=None: _Model(rp))
    _engine._label_cache.clear()
    try:
        vismaps, visclus, save = os.path.join(root, "vm"), os.path.join(root, "vc"), os.path.join(root, "seg")
        a = cotracker_occlusions.extract_object_visibility_data(vdir, mdir, os.path.join(root, "videos"), vismaps, False)
        assert np.array_equal(cotracker_occlusions.load_masks(mdir)[..., 0].numpy(), labels.astype(np.int64))
        if a is None:
            assert g["visibility_rows"] in (None, [])
            return
        rows = [(fr["frame_id"], o["object_id"], o["visibility"]) for fr in a["video_data"] for o in fr["data"]]
        assert len(rows) == len(g["visibility_rows"])
        for (f, o, v), ref in zip(rows, g["visibility_rows"]):
            assert (f, o) == (ref["frame_id"], ref["object_id"])
            assert np.array_equal(np.asarray(v, np.float32), np.asarray(ref["visibility"], np.float32), equal_nan=True)
        with open(os.path.join(vismaps, "ytvis2021", "train", "data", "vid0.json")) as f:
            assert json.load(f)["video_data"] == json.loads(json.dumps(a["video_data"]))

        w = identify_visibility_windows.get_visibility_windows_for_video(a, "ytvis2021", "train", "vid0", visclus,
                                                                         g["visibility_threshold"], False)
        assert json.loads(json.dumps(w["clusters"])) == g["clusters"]
        imgs, imgs_orig, lbls, meta = crw_utils.load_frames_and_masks(vdir, mdir, w, "ytvis2021")
        cm = keymask_utils.save_segmentation_masks(imgs, imgs_orig, lbls, meta, save, False)
        cands = sorted(os.path.relpath(p, cm) for p in glob.glob(os.path.join(cm, "cluster_*", "*.png")))
        assert cands == g["candidate_files"]
        try:
            status = cotracker_matching.temporal_correspondence_match(vdir, mdir, cm, vismaps, visclus,
                                                                      g["matching_threshold"], False)
        except Exception as e:  # noqa: BLE001
            status = f"exception:{type(e).__name__}"
        assert status == g["status"], (status, g["status"])
        nm = len(g["visibility_rows"])
        if g["status"] == 1:
            assert rp.calls == [[c[0], c[1], c[2], bool(c[3])] for c in g["tracker_calls"]]
            groups = sorted(os.path.relpath(p, cm) for p in glob.glob(os.path.join(cm, "cluster_*", "group_*", "*.png")))
            assert groups == g["group_files"]
            with open(os.path.join(cm, "video_coverage.txt")) as f:
                assert f.read() == g["video_coverage_txt"]
            for cname, txt in g["cluster_coverage_txt"].items():
                with open(os.path.join(cm, cname, "cluster_coverage.txt")) as f:
                    assert f.read() == txt
            with open(os.path.join(cm, "video_one2x_data.json")) as f:
                mine = json.load(f)
            assert list(mine) == list(g["one2x"])
            for ck, ref in g["one2x"].items():
                assert list(mine[ck]) == list(ref), (list(mine[ck]), list(ref))
                for k, v in ref.items():
                    if k == "avg_one2x_cluster":
                        assert _close(float(mine[ck][k]), float(v))
                    else:
                        assert _close(float(mine[ck][k]["avg_one2x"]), float(v["avg_one2x"]))
                        assert mine[ck][k]["one2x_counts"] == v["one2x_counts"] and mine[ck][k]["noisy"] == v["noisy"]
            # candidate PNG content: 0/255 masks of the right label
            import cv2
            p0 = os.path.join(cm, g["candidate_files"][0])
            parts = os.path.basename(p0).split("_")
            f0, m0 = int(parts[1][5:]), int(parts[2].split(".")[0][4:])
            assert np.array_equal(cv2.imread(p0, cv2.IMREAD_UNCHANGED), (labels[f0] == m0).astype(np.uint8) * 255)
        else:
            assert len(rp.calls) >= nm
    finally:
        _engine.set_tracker_factory(None)


@pytest.mark.gpu
def test_dropin_helpers_on_gpu():
    from s2d_b200.keymask_ident import cotracker_matching as cm, identify_visibility_windows as ivw
    rng = np.random.default_rng(0)
    H, W, T, P = 40, 60, 3, 200
    tracks = torch.from_numpy(rng.uniform(-5, 65, size=(1, T, P, 2)).astype(np.float32))
    planes = cm.pred_tracks_to_binary_masks(tracks, H, W)
    from oracle import keymask_oracle as ko
    for t in range(T):
        assert np.array_equal(planes[0, t].numpy(), ko.rasterise_tracks(tracks[0, t].numpy(), H, W))
    mask = torch.from_numpy((rng.random((H, W)) < 0.4).astype(np.uint8) * 255)
    iou = cm.compute_point_mask_intersection(planes[0, 0], mask, 25)
    pm = planes[0, 0].numpy() > 0
    assert iou == (pm & (mask.numpy() > 0)).sum() / pm.sum()
    maj = torch.tensor([0, 1, 1, 0, 0, 1, 0, 1, 1, 1], dtype=torch.float32)
    assert ivw.get_visible_ranges(maj) == [(1, 2), (5, 5), (7, 9)]
    assert ivw.get_visible_ranges(torch.zeros(7)) == []
    assert ivw.get_visible_ranges(torch.ones(40)) == [(0, 39)]


@pytest.mark.gpu
@pytest.mark.parametrize("H,W", [(1, 1), (31, 33), (64, 96), (97, 131), (480, 854), (720, 1280)])
def test_rle_encode_gpu_vs_restatement(H, W):
    """f2: s2d_rle_encode (column-major runs, area, bbox) against the numpy restatement of maskApi.c and a
    decode round trip; masks that start set, empty / full masks, ragged sizes."""
    import numpy as np
    from oracle import coco_rle as cr
    from s2d_b200.keymask_ident.annotations import coco_rle
    rng = np.random.default_rng(H * 7 + W)
    yy, xx = np.mgrid[0:H, 0:W]
    masks = [np.zeros((H, W), bool), np.ones((H, W), bool), rng.random((H, W)) < 0.5,
             ((xx - W / 2) ** 2 / max(1, (W / 3) ** 2) + (yy - H / 2) ** 2 / max(1, (H / 4) ** 2)) <= 1.0,
             (xx + yy) % 7 < 3]
    m0 = np.zeros((H, W), bool); m0[0, 0] = True; masks.append(m0)
    m1 = np.zeros((H, W), bool); m1[-1, -1] = True; masks.append(m1)
    enc = coco_rle.encode_batch(np.stack(masks))
    for m, (rle, area, box) in zip(masks, enc):
        assert rle["size"] == [H, W]
        c = coco_rle.from_string(rle["counts"])
        assert c == cr.counts(m)
        assert np.array_equal(cr.decode(c, H, W), m.astype(np.uint8))
        assert area == cr.area(m) == int(m.sum()) and box == cr.bbox(m)


@pytest.mark.gpu
def test_write_annotation_for_video_roundtrip(tmp_path):
    """stage E drop-in on a small cluster/group tree: YTVIS schema of the reference (annotations.py:28-140),
    segmentations decode back to the PNG masks, areas / boxes match."""
    import json
    import numpy as np
    from PIL import Image
    from oracle import coco_rle as cr
    from s2d_b200.keymask_ident import annotations as ann
    H, W, T = 48, 70, 5
    vdir = tmp_path / "ytvis2021" / "train" / "JPEGImages" / "vidA"
    vdir.mkdir(parents=True)
    for t in range(T):
        Image.fromarray(np.zeros((H, W, 3), np.uint8)).save(vdir / f"{t:05d}.jpg")
    cdir = tmp_path / "seg" / "vidA"
    rng = np.random.default_rng(0)
    truth = {}
    for cname, groups in (("cluster_0", ("group_0", "group_1")), ("cluster_1", ("group_0",))):
        for g in groups:
            (cdir / cname / g).mkdir(parents=True)
            for t in rng.choice(T, size=3, replace=False):
                m = np.zeros((H, W), np.uint8)
                y0, x0 = rng.integers(0, H - 10), rng.integers(0, W - 10)
                m[y0:y0 + rng.integers(2, 10), x0:x0 + rng.integers(2, 10)] = 255
                Image.fromarray(m).save(cdir / cname / g / f"frame{t}_mask{int(rng.integers(1, 5))}.png")
                truth[(cname, g, int(t))] = m > 0
        Image.fromarray(np.zeros((H, W), np.uint8)).save(cdir / cname / f"{cname.replace('_', '')}_frame0_mask1.png")
    one2x = {"cluster_0": {"avg_one2x_cluster": 0.1, "group_0": {"avg_one2x": 0.126, "one2x_counts": 3, "noisy": False},
                           "group_1": {"avg_one2x": 0.0, "one2x_counts": 3, "noisy": False}},
             "cluster_1": {"avg_one2x_cluster": 0.0, "group_0": {"avg_one2x": 1.0, "one2x_counts": 3, "noisy": True}}}
    (cdir / "video_one2x_data.json").write_text(json.dumps(one2x))
    vis = {"video_name": "vidA", "clusters": [{"cluster_id": 0, "ranges": [[0, 2]]}, {"cluster_id": 1, "ranges": [[3, 4]]}]}
    out = tmp_path / "ann"
    ann.write_annotation_for_video(str(vdir), str(cdir), str(out), vis)
    d = json.loads((out / "vidA.json").read_text())
    assert d["videos"][0]["height"] == H and d["videos"][0]["width"] == W and d["videos"][0]["length"] == T
    assert d["categories"] == [{"supercategory": "object", "id": 1, "name": "fg"}]
    assert [a["id"] for a in d["annotations"]] == [1, 2, 3]
    order = [("cluster_0", "group_0"), ("cluster_0", "group_1"), ("cluster_1", "group_0")]
    for a, (cname, g) in zip(d["annotations"], order):
        assert a["one2x"] == round(one2x[cname][g]["avg_one2x"], 2)
        assert a["visibility_ranges"] == vis["clusters"][int(cname[-1])]["ranges"]
        for t in range(T):
            m = truth.get((cname, g, t))
            if m is None:
                assert a["segmentations"][t] is None and a["bboxes"][t] is None and a["areas"][t] is None
            else:
                seg = a["segmentations"][t]
                assert np.array_equal(cr.decode(ann.coco_rle.from_string(seg["counts"]), H, W), m.astype(np.uint8))
                assert a["areas"][t] == int(m.sum()) and a["bboxes"][t] == cr.bbox(m)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["basic", "one2x", "medium", "dups"])
def test_exported_stage_d_helpers_match_reference_golden(case, tmp_path):
    """The single-call helpers the reference exports next to temporal_correspondence_match
    (extract_mask_matches, temporal_correspondance_clustering, calculate_cluster_coverage,
    gather_and_save_one2x_data; cotracker_matching.py:665-921) against the intermediates the unmodified
    reference produced for the golden videos."""
    import json
    import numpy as np
    from s2d_b200.keymask_ident import cotracker_matching as cm
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    g = json.load(open(os.path.join(gdir, f"{case}.json")))
    z = np.load(os.path.join(gdir, f"{case}.npz"))
    assert g["status"] == 1
    labels, tracks = z["labels"], z["tracks"]
    T, H, W = labels.shape
    masks = torch.from_numpy(labels.astype(np.int64))[..., None]
    glookup = cm.contruct_frameid_maskid_lookup(masks)
    ncl = max(q["cluster_id"] for q in g["queries"]) + 1
    cluster_masks = [[{"frame_id": q["frame_id"], "mask_id": q["mask_id"], "vis_cluster_id": c}
                      for q in g["queries"] if q["cluster_id"] == c] for c in range(ncl)]
    clookup = cm.contruct_frameid_maskid_cluster_lookup(cluster_masks)
    matches_data = []
    for q in g["queries"]:
        gid = q["overall_mask_id"]
        seg = torch.from_numpy((labels[q["frame_id"]] == q["mask_id"]).astype(np.uint8) * 255)
        m, comps = cm.extract_mask_matches(seg, torch.from_numpy(tracks[gid])[None], masks, q["frame_id"], tuple(q["v_range"]),
                                           q["grid_size"], glookup, clookup, q["cluster_id"], g["matching_threshold"])
        assert [x["overall_mask_id"] for x in m] == q["matches"], (case, gid)
        assert [[c["frame_id"], c["mask_id"], c["overall_mask_id"]] for c in comps] == [c[:3] for c in q["comps"]]
        assert [c["iou"] for c in comps] == [c[5] for c in q["comps"]]                  # python floats: exact
        matches_data.append({"cluster_id": q["cluster_id"], "frame_id": q["frame_id"], "mask_id": q["mask_id"],
                             "overall_mask_id": gid, "one2x": q["one2x"], "matches": m})
    cids, groupings = cm.temporal_correspondance_clustering(matches_data, glookup, False)
    assert cids == [x["cluster_id"] for x in g["groupings"]]
    for got, want in zip(groupings, g["groupings"]):
        assert got["visibility_to_temporal_factor"] == want["factor"]
        assert {str(k): [list(fm) for fm in v] for k, v in got["overall_mask_ids_per_label"].items()} == want["groups"]
    cov, ccov = cm.calculate_cluster_coverage(cluster_masks, groupings)
    assert f"Video Coverage: {cov:.2f}\n" == g["video_coverage_txt"]
    for c, v in zip(cids, ccov):
        assert g["cluster_coverage_txt"][f"cluster_{c}"].startswith(f"Cluster {c} Coverage: {v:.2f}\n")
    for c in cids:
        os.makedirs(tmp_path / f"cluster_{c}")
    cm.gather_and_save_one2x_data(matches_data, groupings, str(tmp_path))
    assert json.load(open(tmp_path / "video_one2x_data.json")) == g["one2x"]


@pytest.mark.gpu
def test_point_grid_helpers_vs_reference_golden():
    """get_points_on_a_grid / extend_pointgrid / compute_point_mask_iou (cotracker_matching.py:506-637, exported,
    never called by the driver) against outputs of the unmodified reference (oracle/make_golden_events.py)."""
    import numpy as np
    from s2d_b200.keymask_ident import cotracker_matching as cm
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pointgrid.npz"))
    for i in range(int(g["n"])):
        pm, mk, gs = torch.from_numpy(g[f"pm{i}"]), torch.from_numpy(g[f"mask{i}"]), int(g[f"grid{i}"])
        assert np.array_equal(cm.extend_pointgrid(pm.bool(), gs).numpy(), g[f"ext{i}"])
        assert cm.compute_point_mask_iou(pm, mk, gs) == float(g[f"iou{i}"])
    assert torch.equal(cm.get_points_on_a_grid(7, (48, 64)), torch.from_numpy(g["grid7"]))


@pytest.mark.gpu
def test_rle_area_bbox_and_results_conversion(tmp_path):
    """s2d_rle_area_bbox against the loop-for-loop restatement of maskApi.c rleArea / rleToBbox, and the
    convert_results_to_annotations drop-in end to end (schema of convert_results_to_annotations.py:38-95)."""
    import json
    import numpy as np
    from oracle import coco_rle as cr
    from s2d_b200.keymask_ident import annotations as ann
    from s2d_b200.keymask_ident.convert_results_to_annotations import convert_results_to_annotation
    rng = np.random.default_rng(4)
    rles, want = [], []
    for (h, w) in [(1, 1), (7, 5), (33, 64), (97, 131), (480, 854)]:
        yy, xx = np.mgrid[0:h, 0:w]
        for m in (np.zeros((h, w), bool), np.ones((h, w), bool), rng.random((h, w)) < 0.3,
                  (xx > w // 3) & (xx < 2 * w // 3 + 1) & (yy >= h // 4) & (yy <= 3 * h // 4), (xx + 2 * yy) % 11 < 4):
            c = cr.counts(m)
            rles.append({"size": [h, w], "counts": ann.coco_rle.to_string(c)})
            want.append((cr.rle_area(c), cr.rle_to_bbox(c, h)))
            if m.any():
                assert cr.rle_to_bbox(c, h) == cr.bbox(m) and cr.rle_area(c) == int(m.sum())
    rles.append({"size": [4, 4], "counts": [3, 5, 8]})                 # a run that wraps over columns: all rows
    want.append((5, cr.rle_to_bbox([3, 5, 8], 4)))
    assert ann.rle_area_bbox_batch(rles) == want
    # end to end
    gt = {"info": {"d": 1}, "licenses": [], "videos": [{"id": 7, "length": 3, "height": 33, "width": 64, "file_names": ["v7/0.jpg"]}]}
    merged = {"categories": [{"id": 1, "name": "fg"}]}
    seg = [rles[10], None, rles[12]]
    results = [{"video_id": 7, "score": 0.9, "category_id": 1, "segmentations": seg},
               {"video_id": 7, "score": 0.1, "category_id": 1, "segmentations": seg}]
    for name, obj in (("gt", gt), ("merged", merged), ("res", results)):
        (tmp_path / f"{name}.json").write_text(json.dumps(obj))
    convert_results_to_annotation(str(tmp_path / "merged.json"), str(tmp_path / "gt.json"), str(tmp_path / "res.json"), 0.75,
                                  str(tmp_path / "out"), "conv")
    d = json.loads((tmp_path / "out" / "conv.json").read_text())
    assert d["videos"] == gt["videos"] and d["categories"] == merged["categories"] and len(d["annotations"]) == 1
    a = d["annotations"][0]
    assert a["id"] == 1 and a["length"] == 3 and a["segmentations"] == seg
    assert a["areas"] == [want[10][0], None, want[12][0]] and a["bboxes"] == [want[10][1], None, want[12][1]]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["basic", "long"])
def test_flat_imports_like_the_reference_driver(case, tmp_path):
    """INTEGRATION.md mode 1: the reference driver imports its stage modules FLAT (main_keymask_ident.py:4-9:
    `import crw_utils`, `from cotracker_occlusions import ...`). With s2d_b200/keymask_ident first on PYTHONPATH those
    names must resolve to the CUDA-backed modules (their `except ImportError` branches), run the five stages in the
    driver's order in a fresh interpreter, and write what the unmodified reference wrote (golden)."""
    import subprocess
    import sys
    import textwrap
    from tests.conftest import load_golden
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g, labels, tracks, vis = load_golden(case)
    vdir, mdir = _write_video(str(tmp_path), labels)
    script = tmp_path / "flat_driver.py"
    script.write_text(textwrap.dedent(f"""
        import glob, json, os, sys
        import numpy as np
        # --- the reference driver's own import block (main_keymask_ident.py:4-9) ---
        import crw_utils
        from cotracker_occlusions import extract_object_visibility_data
        from identify_visibility_windows import get_visibility_windows_for_video
        from keymask_utils import save_segmentation_masks
        from cotracker_matching import temporal_correspondence_match
        from annotations import write_annotation_for_video
        import _engine, cotracker_matching
        here = os.path.realpath(os.path.dirname(crw_utils.__file__))
        assert here == os.path.realpath({os.path.join(root, 's2d_b200', 'keymask_ident')!r}), here
        assert "s2d_b200.keymask_ident.crw_utils" not in sys.modules and cotracker_matching.crw_utils is crw_utils
        sys.path.insert(0, {root!r})
        from tests.test_dropin_gpu import _Model, _Replay
        z = np.load({os.path.join(root, 'tests', 'golden', case + '.npz')!r})
        rp = _Replay(z["labels"], z["tracks"], z["vis"])
        _engine.set_tracker_factory(lambda checkpoint=None: _Model(rp))
        root = {str(tmp_path)!r}
        vismaps, visclus, save = os.path.join(root, "vm"), os.path.join(root, "vc"), os.path.join(root, "seg")
        a = extract_object_visibility_data({vdir!r}, {mdir!r}, os.path.join(root, "videos"), vismaps, False)
        w = get_visibility_windows_for_video(a, "ytvis2021", "train", "vid0", visclus, {g['visibility_threshold']!r}, False)
        imgs, imgs_orig, lbls, meta = crw_utils.load_frames_and_masks({vdir!r}, {mdir!r}, w, "ytvis2021")
        cm = save_segmentation_masks(imgs, imgs_orig, lbls, meta, save, False)
        status = temporal_correspondence_match({vdir!r}, {mdir!r}, cm, vismaps, visclus, {g['matching_threshold']!r}, False)
        write_annotation_for_video({vdir!r}, cm, os.path.join(root, "ann"), w)
        groups = sorted(os.path.relpath(p, cm) for p in glob.glob(os.path.join(cm, "cluster_*", "group_*", "*.png")))
        ann = json.load(open(os.path.join(root, "ann", "vid0.json")))
        json.dump({{"status": status, "clusters": w["clusters"], "groups": groups, "n_ann": len(ann["annotations"]),
                   "length": ann["videos"][0]["length"]}}, open(os.path.join(root, "out.json"), "w"))
    """))
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(root, "s2d_b200", "keymask_ident"), root])
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.load(open(tmp_path / "out.json"))
    assert out["status"] == g["status"] == 1
    assert out["clusters"] == g["clusters"]
    assert out["groups"] == g["group_files"]
    assert out["n_ann"] == len({os.path.dirname(p) for p in g["group_files"]}) and out["length"] == labels.shape[0]


def test_hull_rasters_vs_reference_golden():
    """pred_tracks_to_binary_masks(return_mask=True) (cotracker_matching.py:488-499: convex hull fill, 1-2 point disc
    fallback, empty frames) against rasters produced by the unmodified reference. Host-only (OpenCV), runs without a GPU."""
    from s2d_b200.keymask_ident import cotracker_matching as cm
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pointgrid.npz"))
    H, W = (int(v) for v in g["hull_hw"])
    got = cm.pred_tracks_to_binary_masks(torch.from_numpy(g["hull_tracks"]), H, W, return_mask=True)
    assert got.dtype == torch.uint8 and np.array_equal(got.numpy(), g["hull_masks"])
    assert g["hull_masks"][1, 2].sum() == 0 and 0 < g["hull_masks"][1, 4].sum() <= 5 and g["hull_masks"][0, 2].sum() > 1000


@pytest.mark.gpu
def test_point_rasters_vs_reference_golden():
    """pred_tracks_to_binary_masks(return_mask=False) - the rasterise kernel - against the unmodified reference's rasters."""
    from s2d_b200.keymask_ident import cotracker_matching as cm
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pointgrid.npz"))
    H, W = (int(v) for v in g["hull_hw"])
    got = cm.pred_tracks_to_binary_masks(torch.from_numpy(g["hull_tracks"]), H, W)
    assert np.array_equal(got.cpu().numpy(), g["point_masks"])
