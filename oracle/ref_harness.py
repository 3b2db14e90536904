"""TEST INFRASTRUCTURE ONLY - never imported by the product path (s2d_b200/).

Runs the *unmodified* reference (`/root/reference/keymask_ident`) end to end on a synthetic
on-disk video and snapshots every intermediate of the hot path. It exists to (i) pin the CPU
restatement in `oracle/keymask_oracle.py` and (ii) generate the committed fixtures under
`tests/golden/` (see `oracle/make_golden.py`). It only works inside the build container,
where `/root/reference` is mounted; nothing that runs on the GPU box may import it.

Recipe: stub the absent third-party modules (`cotracker`,
`matplotlib`, `imageio`), inject a FakePredictor that replays the scene's tracks/visibility
for whichever (frame, mask) the reference asks for, pretend the checkpoint dir exists, lay the
video out under a path containing `ytvis2021/train`, then call the five stage functions in
driver order (main_keymask_ident.py:91-128).
"""
from __future__ import annotations

import contextlib
import glob
import io
import json
import os
import sys
import time
import types

import numpy as np

REFERENCE_DIR = "/root/reference/keymask_ident"


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_DIR)


class _FakePredictorFactory:
    """Builds a CoTrackerPredictor stand-in bound to one scene."""

    def __init__(self, scene):
        self.scene = scene
        self.lookup = {(int(f), int(l)): q for q, (f, l) in
                       enumerate(zip(scene.query_frame, scene.query_label))}
        self.calls = []

    def __call__(self, checkpoint=None):
        factory = self
        import torch

        class FakePredictor:
            def cuda(self):
                return self

            def __call__(self, video, grid_size=0, grid_query_frame=0, segm_mask=None,
                         backward_tracking=False):
                sm = segm_mask[0, 0].numpy()
                ys, xs = np.nonzero(sm)
                lab = int(factory.scene.labels[grid_query_frame][ys[0], xs[0]])
                q = factory.lookup[(int(grid_query_frame), lab)]
                factory.calls.append((int(grid_query_frame), lab, int(grid_size), bool(backward_tracking)))
                tr = torch.from_numpy(factory.scene.tracks[q][None].copy())
                vs = torch.from_numpy(factory.scene.vis[q][None].astype(bool))
                return tr, vs

        return FakePredictor()


def _install_stubs(predictor_factory):
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("cotracker")
    stub("cotracker.predictor", CoTrackerPredictor=predictor_factory)
    stub("cotracker.utils")
    stub("cotracker.utils.visualizer", Visualizer=object, read_video_from_path=None)
    if "matplotlib" not in sys.modules:
        mpl = stub("matplotlib")
        mpl.pyplot = stub("matplotlib.pyplot")
        mpl.cm = stub("matplotlib.cm")
    if "imageio" not in sys.modules:
        stub("imageio")


def write_scene_to_disk(scene, root, video_name="vid0"):
    """frames  <root>/ytvis2021/train/JPEGImages/<vid>/00000.jpg ...
       masks   <root>/masks/<vid>/00000.png  colour coded, black background."""
    import cv2
    from s2d_b200.synth import scene_object_map

    T, H, W = scene.labels.shape
    vdir = os.path.join(root, "ytvis2021", "train", "JPEGImages", video_name)
    mdir = os.path.join(root, "masks", video_name)
    os.makedirs(vdir, exist_ok=True)
    os.makedirs(mdir, exist_ok=True)
    omap = scene_object_map(scene)
    for t in range(T):
        rgb = np.zeros((H, W, 3), np.uint8)
        m = omap[t] >= 0
        rgb[m] = scene.colors[omap[t][m]]
        cv2.imwrite(os.path.join(mdir, f"{t:05d}.png"), rgb[..., ::-1])
        frame = np.full((H, W, 3), 127, np.uint8)
        cv2.imwrite(os.path.join(vdir, f"{t:05d}.jpg"), frame)
    return vdir, mdir


def run_reference(scene, workdir, *, visibility_threshold=0.3, matching_threshold=0.5,
                  video_name="vid0", quiet=True, log_pairs=True):
    """Returns a JSON-able dict with every intermediate of stages A-D. `log_pairs=False` leaves
    compute_point_mask_intersection unwrapped (no per-pair (I, U) record): the timing runs use it so that the
    harness adds nothing to the reference's inner loop."""
    factory = _FakePredictorFactory(scene)
    _install_stubs(factory)
    real_exists = os.path.exists
    os.path.exists = lambda p: True if str(p).rstrip("/") == "/mnt/data/checkpoints" else real_exists(p)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    for name in ("crw_utils", "cotracker_occlusions", "identify_visibility_windows",
                 "keymask_utils", "cotracker_matching"):
        sys.modules.pop(name, None)
    try:
        import crw_utils, cotracker_occlusions, identify_visibility_windows  # noqa: E401
        import keymask_utils, cotracker_matching                           # noqa: E401
        cotracker_occlusions.CoTrackerPredictor = factory
        cotracker_matching.CoTrackerPredictor = factory

        vdir, mdir = write_scene_to_disk(scene, workdir, video_name)
        save_path = os.path.join(workdir, "seg")
        vismaps = os.path.join(workdir, "vismaps")
        visclus = os.path.join(workdir, "visclusters")

        # record per-pair integer counts next to the reference's float
        pair_log = []
        orig_pmi = cotracker_matching.compute_point_mask_intersection

        def logged_pmi(pointmask, mask, grid_size):
            import torch
            iou = orig_pmi(pointmask, mask, grid_size)
            pm = pointmask.bool()
            mk = mask.bool() * pm
            pair_log.append((int(torch.sum(pm & mk).item()), int(torch.sum(pm | mk).item()), float(iou)))
            return iou

        if log_pairs:
            cotracker_matching.compute_point_mask_intersection = logged_pmi
        comparisons_log = []
        orig_emm = cotracker_matching.extract_mask_matches

        timing = {"extract_mask_matches_s": 0.0, "extract_mask_matches_calls": 0}

        def logged_emm(*a, **k):
            start = len(pair_log)
            _t = time.perf_counter()
            matches, comps = orig_emm(*a, **k)
            timing["extract_mask_matches_s"] += time.perf_counter() - _t
            timing["extract_mask_matches_calls"] += 1
            comparisons_log.append(dict(frame_id=int(a[3]), v_range=[int(a[4][0]), int(a[4][1])],
                                        grid_size=int(a[5]), pairs=pair_log[start:],
                                        comps=[(int(c["frame_id"]), int(c["mask_id"]),
                                                int(c["overall_mask_id"]), float(c["iou"])) for c in comps],
                                        matches=[int(m["overall_mask_id"]) for m in matches]))
            return matches, comps

        cotracker_matching.extract_mask_matches = logged_emm
        clustering_log = {}
        orig_tcc = cotracker_matching.temporal_correspondance_clustering

        def logged_tcc(matches_data, lookup, debug):
            clustering_log["matches_data"] = [
                dict(cluster_id=int(m["cluster_id"]), frame_id=int(m["frame_id"]), mask_id=int(m["mask_id"]),
                     overall_mask_id=int(m["overall_mask_id"]), one2x=int(m["one2x"]),
                     matches=[int(x["overall_mask_id"]) for x in m["matches"]]) for m in matches_data]
            out = orig_tcc(matches_data, lookup, debug)
            clustering_log["out"] = out
            return out

        cotracker_matching.temporal_correspondance_clustering = logged_tcc

        sink = io.StringIO()
        ctx = contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()
        ctx2 = contextlib.redirect_stderr(sink) if quiet else contextlib.nullcontext()
        import warnings
        with ctx, ctx2, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = {"timing": timing}
            _t0 = time.perf_counter()
            visdata = cotracker_occlusions.extract_object_visibility_data(vdir, mdir, os.path.join(workdir, "videos"),
                                                                          vismaps, False)
            timing["stage_a_s"] = time.perf_counter() - _t0
            out["stage_a"] = visdata
            ref_labels = cotracker_occlusions.load_masks(mdir)[..., 0].numpy()
            out["labels_equal"] = bool(np.array_equal(ref_labels, scene.labels.astype(np.int64)))
            if visdata is None:
                out["status"] = None
                return out
            _t0 = time.perf_counter()
            windows = identify_visibility_windows.get_visibility_windows_for_video(
                visdata, "ytvis2021", "train", video_name, visclus, visibility_threshold, False)
            timing["stage_b_s"] = time.perf_counter() - _t0
            out["stage_b"] = json.loads(json.dumps(windows))
            _t0 = time.perf_counter()
            imgs, imgs_orig, lbls, meta = crw_utils.load_frames_and_masks(vdir, mdir, windows, "ytvis2021")
            cm_path = keymask_utils.save_segmentation_masks(imgs, imgs_orig, lbls, meta, save_path, False)
            timing["stage_c_s"] = time.perf_counter() - _t0
            out["candidate_files"] = sorted(os.path.relpath(p, cm_path) for p in
                                            glob.glob(os.path.join(cm_path, "cluster_*", "*.png")))
            _t0 = time.perf_counter()
            try:
                status = cotracker_matching.temporal_correspondence_match(
                    vdir, mdir, cm_path, vismaps, visclus, matching_threshold, False)
            except Exception as e:  # the driver catches per stage (main_keymask_ident.py:127-132)
                status = f"exception:{type(e).__name__}"
            timing["stage_d_s"] = time.perf_counter() - _t0
            out["status"] = status
            out["tracker_calls"] = factory.calls
            out["queries"] = comparisons_log
            out["matches_data"] = clustering_log.get("matches_data")
            tcc = clustering_log.get("out")
            if tcc is not None and tcc[0] != -1:
                out["groupings"] = [dict(cluster_id=int(g["cluster_id"]),
                                         factor=int(g["visibility_to_temporal_factor"]),
                                         groups={str(k): [[int(f), int(m)] for f, m in v]
                                                 for k, v in g["overall_mask_ids_per_label"].items()})
                                    for g in tcc[1]]
            else:
                out["groupings"] = None
            if status == 1:
                out["group_files"] = sorted(os.path.relpath(p, cm_path) for p in
                                            glob.glob(os.path.join(cm_path, "cluster_*", "group_*", "*.png")))
                with open(os.path.join(cm_path, "video_one2x_data.json")) as f:
                    out["one2x"] = json.load(f)
                with open(os.path.join(cm_path, "video_coverage.txt")) as f:
                    out["video_coverage_txt"] = f.read()
                out["cluster_coverage_txt"] = {}
                for p in sorted(glob.glob(os.path.join(cm_path, "cluster_*", "cluster_coverage.txt"))):
                    with open(p) as f:
                        out["cluster_coverage_txt"][os.path.basename(os.path.dirname(p))] = f.read()
        return out
    finally:
        os.path.exists = real_exists
