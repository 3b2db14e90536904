"""Drop-in for keymask_ident/cotracker_matching.py (stage D): point-in-mask voting (K2),
scoring/selection (K4a) and temporal-correspondence grouping (K4b) run on the GPU; this module
keeps the reference's file protocol, ordering rules and failure sentinels."""
from __future__ import annotations

import glob
import json
import os
import shutil
import warnings

import cv2
import numpy as np
import torch
from PIL import Image

try:
    from . import _engine, crw_utils
    from .cotracker_occlusions import boolean_visibility, get_segmentation_mask as _seg_thw1, mp4_from_images
except ImportError:
    import _engine
    import crw_utils
    from cotracker_occlusions import boolean_visibility, get_segmentation_mask as _seg_thw1, mp4_from_images

from s2d_b200 import _lib
from s2d_b200.pipeline import Batch, Params, VideoInput

load_image_robust = crw_utils.load_image_robust


def load_masks(mask_folder: str):
    if not sorted(glob.glob(os.path.join(mask_folder, "*.png"))):
        warnings.warn(f"No .png masks found in {mask_folder!r}")
        return None
    return crw_utils.load_masks(mask_folder)


def load_cluster_masks(mask_folder: str):
    """non-empty `cluster_*` folders in lexicographic order, masks in lexicographic file order
    (cotracker_matching.py:87-128)."""
    folders = [f for f in sorted(glob.glob(os.path.join(mask_folder, "cluster_*"))) if len(os.listdir(f)) > 0]
    if not folders:
        warnings.warn(f"No cluster folders found in {mask_folder!r}. Skipping video!")
        return []
    out = []
    for folder in folders:
        cid = int(os.path.basename(folder).split("_")[1])
        entries = []
        for path in sorted(glob.glob(os.path.join(folder, "*.png"))):
            parts = os.path.basename(path).split("_")
            mask = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            if mask is None:
                continue
            entries.append({"vis_cluster_id": cid, "frame_id": int(parts[1].replace("frame", "")),
                            "mask_id": int(parts[2].split(".")[0].replace("mask", "")),
                            "mask": (mask > 0).astype(np.uint8) * 255})
        out.append(entries)
    return out


def get_segmentation_mask(masks: torch.Tensor, query_frame_idx: int, object_id: int = 1) -> torch.Tensor:
    if query_frame_idx >= 0:
        return _seg_thw1(masks, query_frame_idx, object_id)
    frame = masks[..., 0]
    sel = (frame != 0) if object_id == -1 else (frame == object_id)
    return (sel.to(torch.uint8) * 255)[None, None]


def contruct_frameid_maskid_lookup(all_video_masks):
    """global ids in (frame, label) order over sort(unique(label[t]))[1:] - enumerated by K0."""
    qf, ql, _ = _engine.enumerate_objects(_engine.labels_u8_device(all_video_masks))
    out = [[] for _ in range(all_video_masks.shape[0])]
    for g, (f, l) in enumerate(zip(qf.tolist(), ql.tolist())):
        out[f].append({"frame_id": f, "mask_id": l, "overall_mask_id": g})
    return out


def contruct_frameid_maskid_cluster_lookup(cluster_masks):
    return [[{"cluster_mask_id": i, "frame_id": m["frame_id"], "mask_id": m["mask_id"]} for i, m in enumerate(c)]
            for c in cluster_masks]


def get_overall_maskid(lookup, frame_id, mask_id):
    return next((e["overall_mask_id"] for e in lookup[frame_id] if e["mask_id"] == mask_id), None)


def get_cluster_maskid(lookup, cluster_id, frame_id, mask_id):
    return next((e["cluster_mask_id"] for e in lookup[cluster_id]
                 if e["frame_id"] == frame_id and e["mask_id"] == mask_id), None)


def get_frameid_maskid_from_overall_maskid(lookup, overall_mask_id):
    for frame in lookup:
        for e in frame:
            if e["overall_mask_id"] == overall_mask_id:
                return e["frame_id"], e["mask_id"]
    raise ValueError(f"Overall mask ID {overall_mask_id} not found in lookup table.")


def load_visibility_data(visibility_maps_output_dir, video_name):
    with open(os.path.join(visibility_maps_output_dir, f"{video_name}.json")) as f:
        return json.load(f)["clusters"]


def get_masks_for_vrange(masks, v_range):
    return [m for m in masks if v_range[0] <= m["frame_id"] <= v_range[1]]


def pred_tracks_to_binary_masks(pred_tracks: torch.Tensor, height: int, width: int, return_mask: bool = False):
    """(B,T,P,2) tracks -> (B,T,H,W) uint8 rasters of the rounded in-bounds points (K1's A operand,
    rasterise kernel). The convex-hull variant (return_mask=True) is not on the keymask path."""
    if return_mask:
        raise NotImplementedError("convex-hull rasterisation is not used by keymask discovery")
    B, T, P, _ = pred_tracks.shape
    dev = _engine.device()
    out = torch.empty((B, T, height, width), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    for b in range(B):
        tr = pred_tracks[b].to(dev, torch.float32).contiguous()
        _lib.call("s2d_rasterise_tracks", tr.data_ptr(), T, P, height, width, out[b].data_ptr(), st)
    torch.cuda.synchronize(dev)
    return out.to(pred_tracks.device)


def compute_point_mask_intersection(pointmask: torch.Tensor, mask: torch.Tensor, grid_size: int) -> float:
    """|P and mask| / |P or (mask and P)| through the bit-packed overlap kernel (K1)."""
    dev = _engine.device()
    st = torch.cuda.current_stream(dev).cuda_stream
    a = (pointmask != 0).to(dev, torch.uint8).contiguous().reshape(1, -1)
    b = (mask != 0).to(dev, torch.uint8).contiguous().reshape(1, -1)
    npix = a.shape[1]
    nw = (npix + 31) // 32
    ba = torch.empty(nw, dtype=torch.int32, device=dev)
    bb = torch.empty(nw, dtype=torch.int32, device=dev)
    I = torch.empty(1, dtype=torch.int32, device=dev)
    A = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("s2d_pack_bits", a.data_ptr(), 1, npix, ba.data_ptr(), st)
    _lib.call("s2d_pack_bits", b.data_ptr(), 1, npix, bb.data_ptr(), st)
    _lib.call("s2d_overlap_bits", ba.data_ptr(), 1, bb.data_ptr(), 1, nw, I.data_ptr(), A.data_ptr(), None, st)
    inter, union = int(I.item()), int(A.item())
    return 0.0 if union == 0 else inter / union


def extract_appearance_events(vis: torch.Tensor, smoothing_window: int = 1, thresh: float = 0.95, min_run_length: int = 4):
    """{row: [(appear_frame, disappear_frame), ...]} of (N, T) visibility curves; exported by the
    reference but never called on the keymask path (cotracker_matching.py:212-269); K3d
    appearance_events_kernel."""
    return _engine.appearance_events(vis, smoothing_window, thresh, min_run_length)


def boolean_visibility(vis: torch.Tensor, threshold: float = 0.3) -> torch.Tensor:
    """vis >= threshold on the GPU (cotracker_matching.py:272-286)."""
    return _engine.boolean_visibility(vis, threshold)


def crop_bool_tensor(bool_arr: np.ndarray):
    if bool_arr.ndim != 2:
        raise ValueError("Input must be a 2D boolean array.")
    rows, cols = np.nonzero(np.any(bool_arr, axis=1))[0], np.nonzero(np.any(bool_arr, axis=0))[0]
    if len(rows) == 0 or len(cols) == 0:
        return np.zeros((0, 0), dtype=bool), (0, 0)
    return bool_arr[rows[0]:rows[-1] + 1, cols[0]:cols[-1] + 1], (rows[0], cols[0])


def save_temporal_group_masks(mask_groupings, cluster_masks, visibility_group_mask_path, idx_correction=0):
    for g in mask_groupings:
        cid = g["cluster_id"]
        cdir = os.path.join(visibility_group_mask_path, f"cluster_{cid}")
        for old in glob.glob(os.path.join(cdir, "group_*")):
            shutil.rmtree(old)
        for label, fms in g["overall_mask_ids_per_label"].items():
            gdir = os.path.join(cdir, f"group_{label}")
            os.makedirs(gdir, exist_ok=True)
            masks = cluster_masks[cid - idx_correction] if cid >= len(cluster_masks) else cluster_masks[cid]
            for (f, m) in fms:
                hit = next((x for x in masks if x["frame_id"] == f and x["mask_id"] == m), None)
                if hit is not None:
                    Image.fromarray(hit["mask"]).save(os.path.join(gdir, f"frame{f}_mask{m}.png"))


def save_cluster_coverages(video_coverage, cluster_coverages, visibility_to_temporal_factors, cluster_mask_path):
    with open(os.path.join(cluster_mask_path, "video_coverage.txt"), "w") as f:
        f.write(f"Video Coverage: {video_coverage:.2f}\n")
    cids = sorted(int(d.split("_")[1]) for d in os.listdir(cluster_mask_path)
                  if d.startswith("cluster_") and os.path.isdir(os.path.join(cluster_mask_path, d)))
    for i, cov in enumerate(cluster_coverages):
        with open(os.path.join(cluster_mask_path, f"cluster_{cids[i]}", "cluster_coverage.txt"), "w") as f:
            f.write(f"Cluster {cids[i]} Coverage: {cov:.2f}\n"
                    f"Visibility to Temporal Factor: {visibility_to_temporal_factors[i]}\n")


def _write_one2x(one2x_video, cluster_mask_path):
    for cname, od in one2x_video.items():
        cid = int(cname.split("_")[1])
        with open(os.path.join(cluster_mask_path, cname, f"one2x_data_cluster{cid}.json"), "w") as f:
            json.dump(od, f, indent=4)
    with open(os.path.join(cluster_mask_path, "video_one2x_data.json"), "w") as f:
        json.dump(one2x_video, f, indent=4)


def temporal_correspondence_match(video_path, mask_path, cluster_mask_path, visibility_maps_output_base,
                                  visibility_clusters_output_base, matching_threshold, debug=False):
    """Same contract as the reference (cotracker_matching.py:926-1136): 1 on success, -1 when the
    video cannot be annotated; writes group PNGs, coverage and one2x files."""
    all_video_masks = _engine.cached_labels(mask_path, load_masks)
    if all_video_masks is None:
        print("Failed to load masks for temporal correspondence matching...")
        return -1
    cluster_masks = load_cluster_masks(cluster_mask_path)
    if len(cluster_masks) == 0:
        return -1
    dataset_name, split = _engine.dataset_and_split(video_path)
    os.makedirs(os.path.join(visibility_maps_output_base, dataset_name, split), exist_ok=True)
    video_name = os.path.basename(video_path)
    video = mp4_from_images(video_path)
    clusters = load_visibility_data(os.path.join(visibility_clusters_output_base, dataset_name, split), video_name)
    model = _engine.make_tracker()
    if torch.cuda.is_available():
        video = video.cuda()
    clusters.sort(key=lambda c: int(c["cluster_id"]))
    if len(cluster_masks) != len(clusters):
        warnings.warn(f"Cluster masks length {len(cluster_masks)} does not match visibility ranges length {len(clusters)}")
        return -1

    labels_dev = _engine.labels_u8_device(all_video_masks)
    T, H, W = labels_dev.shape
    qf, ql, area = _engine.enumerate_objects(labels_dev)
    gid = {(int(f), int(l)): g for g, (f, l) in enumerate(zip(qf, ql))}
    nm = len(qf)
    rowinfo = np.full((nm, 4), -1, np.int32)

    # host protocol of the reference: which masks are queries, in which window (file derived)
    order, tracks_of = [], {}
    for c in clusters:
        cid = int(c["cluster_id"])
        if len(c["ranges"]) == 0:
            continue
        v0, v1 = min(r[0] for r in c["ranges"]), max(r[1] for r in c["ranges"])
        cm = cluster_masks[cid]                                        # positional, like the reference
        if len(cm) == 0:
            continue
        if cm[0]["vis_cluster_id"] != cid:
            print(f"Cluster ID mismatch: {cm[0]['vis_cluster_id']} != {cid}")
            return -1
        for md in sorted(get_masks_for_vrange(cm, (v0, v1)), key=lambda x: int(x["frame_id"])):
            f, m = md["frame_id"], md["mask_id"]
            grid = max(min(int(np.sum(md["mask"] / 255) // 800), 50), 25)
            tracks, _vis = model(video, grid_size=grid, grid_query_frame=f,
                                 segm_mask=torch.from_numpy(md["mask"])[None, None], backward_tracking=f > v0)
            g = gid[(f, m)]
            rowinfo[g] = (cid, 0, v0, v1)
            tracks_of[g] = tracks[0]
            order.append((cid, f, m, g))

    if not order:       # nothing to match: the reference ends in an empty crop -> -1
        return -1
    pmax = max(int(t.shape[1]) for t in tracks_of.values())
    pmax += pmax % 2
    dev = labels_dev.device
    tracks = torch.zeros((nm, T, pmax, 2), dtype=torch.float32, device=dev)
    npts = torch.zeros(nm, dtype=torch.int32)
    for g, t in tracks_of.items():
        tracks[g, :, : t.shape[1]] = t.to(dev, torch.float32)
        npts[g] = t.shape[1]
    b = Batch([VideoInput(labels=labels_dev, tracks=tracks, npts=npts.to(dev))], stages="LD")
    b.upload_stage_b(rowinfo, len(clusters), 1)
    b.run(Params(matching_threshold=matching_threshold))
    torch.cuda.synchronize(dev)
    res = b.decode(want_comps=False, check_rows=True)[0]
    if res["status"] != 1:
        return -1
    groupings = res["groupings"]
    save_temporal_group_masks(groupings, cluster_masks, cluster_mask_path, 0)
    save_cluster_coverages(res["video_coverage"], res["cluster_coverages"],
                           [g["visibility_to_temporal_factor"] for g in groupings], cluster_mask_path)
    _write_one2x(res["one2x"], cluster_mask_path)
    return 1
