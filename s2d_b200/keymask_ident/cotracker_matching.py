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
    rasterise kernel). return_mask=True fills the convex hull of each frame's points instead
    (cotracker_matching.py:488-499; not used by the driver): the hull and the polygon fill stay OpenCV's on the
    host, exactly the calls the reference makes, so the rasters are identical."""
    if return_mask:
        return _hull_masks(pred_tracks, height, width)
    B, T, P, _ = pred_tracks.shape
    dev = _engine.device()
    out = torch.empty((B, T, height, width), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    for b in range(B):
        tr = pred_tracks[b].to(dev, torch.float32).contiguous()
        _lib.call("s2d_rasterise_tracks", tr.data_ptr(), T, P, height, width, out[b].data_ptr(), st)
    torch.cuda.synchronize(dev)
    return out.to(pred_tracks.device)


def _hull_masks(pred_tracks: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """convex-hull rasters: >= 3 in-bounds points -> filled hull, 1-2 points -> radius-1 discs, none -> empty."""
    B, T, _, _ = pred_tracks.shape
    xy = torch.round(pred_tracks.detach().to("cpu", torch.float32)).to(torch.int64).numpy()   # half-to-even; NaN / inf -> INT64_MIN
    out = np.zeros((B, T, height, width), np.uint8)
    for b, t in np.ndindex(B, T):
        p = xy[b, t]
        inside = (p[:, 0] >= 0) & (p[:, 0] < width) & (p[:, 1] >= 0) & (p[:, 1] < height)
        p = p[inside]
        if len(p) >= 3:
            cv2.fillPoly(out[b, t], [cv2.convexHull(p.astype(np.int32))], color=1)
        else:
            for x, y in p:
                cv2.circle(out[b, t], (int(x), int(y)), radius=1, color=1, thickness=-1)
    return torch.from_numpy(out).to(pred_tracks.device)


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


def get_points_on_a_grid(size: int, extent, center=None, device=torch.device("cpu")):
    """(1, size*size, 2) grid of (x, y) points covering an (H, W) extent with a margin of W/64, row-major
    (cotracker_matching.py:506-562, CoTracker's query grid)."""
    if size == 1:
        return torch.tensor([extent[1] / 2, extent[0] / 2], device=device)[None, None]
    if center is None:
        center = [extent[0] / 2, extent[1] / 2]
    margin = extent[1] / 64
    ys = torch.linspace(margin - extent[0] / 2 + center[0], extent[0] / 2 + center[0] - margin, size, device=device)
    xs = torch.linspace(margin - extent[1] / 2 + center[1], extent[1] / 2 + center[1] - margin, size, device=device)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).reshape(1, -1, 2)


def extend_pointgrid(pointmask: torch.Tensor, grid_size: int) -> torch.Tensor:
    """Point raster OR the grid points that fall outside the convex hull of the points
    (cotracker_matching.py:565-609; exported, never called by the driver). Hull and polygon fill stay
    OpenCV's, as in the reference; the grid density follows points-per-hull-area."""
    H, W = pointmask.shape
    pts = torch.nonzero(pointmask, as_tuple=False).cpu().numpy()
    if pts.shape[0] < 3:
        warnings.warn("Not enough points to compute a convex hull. At least 3 points are required.")
        return pointmask
    hull = cv2.convexHull(pts.astype(np.int32))
    filled = np.zeros((H, W), dtype=np.uint8)
    cv2.fillPoly(filled, [hull], color=255)
    filled = torch.from_numpy(filled).to(pointmask.device).bool()
    new_size = torch.sqrt((torch.sum(pointmask) / torch.sum(filled)) * H * W).item()
    if not torch.isfinite(torch.Tensor([new_size])) or new_size <= 0:
        new_size = grid_size
    new_size = min(new_size, grid_size * 1.5)
    grid = get_points_on_a_grid(size=int(new_size), extent=(H, W), device=pointmask.device)
    gx, gy = grid[0, :, 0].round().long(), grid[0, :, 1].round().long()
    outside = ~filled[gy.cpu(), gx.cpu()].bool()
    ext = torch.zeros_like(pointmask, dtype=torch.uint8)
    ext[gy[outside], gx[outside]] = 1
    return pointmask | ext


def compute_point_mask_iou(pointmask: torch.Tensor, mask: torch.Tensor, grid_size: int) -> float:
    """|P and M'| / |P or M'| with M' = mask restricted to the hull-extended point grid
    (cotracker_matching.py:612-637); the counts come from the bit-packed overlap kernel (K1)."""
    pm = pointmask.bool().to(mask.device)
    ext = extend_pointgrid(pm, grid_size)
    m = mask.bool() * ext
    dev = _engine.device()
    st = torch.cuda.current_stream(dev).cuda_stream
    a = pm.to(dev, torch.uint8).contiguous().reshape(1, -1)
    b = m.to(dev, torch.uint8).contiguous().reshape(1, -1)
    npix = a.shape[1]
    nw = (npix + 31) // 32
    ba = torch.empty(nw, dtype=torch.int32, device=dev)
    bb = torch.empty(nw, dtype=torch.int32, device=dev)
    I = torch.empty(1, dtype=torch.int32, device=dev)
    A = torch.empty(1, dtype=torch.int32, device=dev)
    B = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("s2d_pack_bits", a.data_ptr(), 1, npix, ba.data_ptr(), st)
    _lib.call("s2d_pack_bits", b.data_ptr(), 1, npix, bb.data_ptr(), st)
    _lib.call("s2d_overlap_bits", ba.data_ptr(), 1, bb.data_ptr(), 1, nw, I.data_ptr(), A.data_ptr(), B.data_ptr(), st)
    inter = int(I.item())
    union = int(A.item()) + int(B.item()) - inter
    return 0.0 if union == 0 else inter / union


def extract_mask_matches(segm_mask, pred_tracks, all_video_masks, frame_id, v_range, grid_size, globalid_lookup,
                         clusterid_lookup, cluster_id, matching_threshold):
    """Single-query form of stage D's inner loop (cotracker_matching.py:665-719): the query's tracked points
    voted into every frame of v_range by K2, iou = |P and mask| / |P| as python floats, matches where
    iou > matching_threshold. Returns (matches, all_comparisons) with the reference's dict keys. The driver
    path (temporal_correspondence_match) runs the same kernel batched over all queries instead."""
    T = all_video_masks.shape[0]
    assert pred_tracks.shape[1] == T, f"Track masks shape {pred_tracks.shape[1]} does not match video masks shape {T}"
    lab = all_video_masks[..., 0]
    if tuple(lab.shape[1:]) != tuple(segm_mask.shape[:2]):      # targets are nearest-resized to the query mask's size
        lab = torch.nn.functional.interpolate(lab[:, None].float(), size=tuple(segm_mask.shape[:2]), mode="nearest")[:, 0].long()
    labels_dev = _engine.labels_u8_device(lab[..., None])
    dev = labels_dev.device
    qf, ql, _area = _engine.enumerate_objects(labels_dev)
    tr = pred_tracks[0].to(dev, torch.float32)
    P = int(tr.shape[1])
    Pp = P + (P % 2)
    tracks = torch.zeros((1, T, Pp, 2), dtype=torch.float32, device=dev)
    tracks[0, :, :P] = tr
    b = Batch([VideoInput(labels=labels_dev, tracks=tracks, npts=torch.tensor([P], dtype=torch.int32, device=dev),
                          max_label=int(lab.max()))], stages="D")
    b.votes_all()
    torch.cuda.synchronize(dev)
    L = b.host_descs[0].L
    hits = b.hits.cpu().numpy().reshape(T, L)
    uniq = b.uniq.cpu().numpy().reshape(T)
    matches, all_comparisons = [], []
    for f in range(v_range[0], v_range[1] + 1):
        for oid in (int(l) for l in ql[qf == f]):
            inter, union = int(hits[f, oid]), int(uniq[f])
            iou = 0.0 if union == 0 else inter / union
            rec = {"frame_id": f, "mask_id": oid,
                   "overall_mask_id": get_overall_maskid(globalid_lookup, f, oid),
                   "cluster_mask_id": get_cluster_maskid(clusterid_lookup, cluster_id, f, oid), "iou": iou}
            all_comparisons.append(rec)
            if iou > matching_threshold:
                matches.append(dict(rec))
    return matches, all_comparisons


def temporal_correspondance_clustering(matches_data, frameid_maskid_to_overall_maskid_lookup, debug):
    """Per visibility cluster: 0/1 match matrix over overall mask ids, bounding-box crop, Hamming DBSCAN with the
    reference's eps / min_samples table (on the GPU: s2d_hamming_dbscan), zero rows -> -1, groups keyed by label
    (cotracker_matching.py:764-840). (-1, -1) when a cluster's crop is empty."""
    import ctypes as C
    max_id = max([m["overall_mask_id"] for md in matches_data for m in md["matches"]], default=-1)
    cluster_ids = sorted(set(m["cluster_id"] for m in matches_data))
    out = []
    for cid in cluster_ids:
        mat = np.zeros((max_id + 1, max_id + 1), dtype=np.float32)
        for md in (m for m in matches_data if int(m["cluster_id"]) == cid):
            r = md["overall_mask_id"]
            for m in md["matches"]:
                c = m["overall_mask_id"]
                if r >= mat.shape[0] or c >= mat.shape[1]:
                    warnings.warn("Overall mask ID exceeds matrix dimensions. Skipping this match.")
                    continue
                mat[r, c] = 1
        mat, (row_off, _col_off) = crop_bool_tensor(mat)
        if mat.shape[0] == 0 or mat.shape[1] == 0:
            return -1, -1
        n, d = mat.shape
        eps, ms = (0.05, 5) if d > 50 else ((0.1, 3) if d < 10 else (0.1, 5))
        dev = _engine.device()
        stride = (d + 31) // 32
        pad = np.zeros((n, stride * 32), bool)
        pad[:, :d] = mat != 0
        words = np.packbits(pad.reshape(n, stride, 32), axis=-1, bitorder="little").view(np.uint32).reshape(n, stride)
        bits = torch.from_numpy(words.view(np.int32).copy()).to(dev)
        nw = C.c_int64()
        _lib.call("s2d_dbscan_work_ints", n, 1, C.byref(nw))
        work = torch.empty(nw.value + 2, dtype=torch.int32, device=dev)
        lab_d = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.call("s2d_hamming_dbscan", bits.data_ptr(), n, stride, d, eps, ms, work.data_ptr(), lab_d.data_ptr(),
                  torch.cuda.current_stream(dev).cuda_stream)
        labels = lab_d.cpu().numpy().astype(np.int64)
        labels[mat.sum(axis=1) == 0] = -1
        print(f"Clustering labels for cluster {cid}: {labels}") if debug else None
        groups = {}
        for i, l in enumerate(labels.tolist()):
            if l != -1:
                groups.setdefault(l, []).append(get_frameid_maskid_from_overall_maskid(frameid_maskid_to_overall_maskid_lookup, i + row_off))
        out.append({"cluster_id": cid, "visibility_to_temporal_factor": len(set(labels[labels != -1].tolist())),
                    "overall_mask_ids_per_label": groups})
    return cluster_ids, out


def calculate_cluster_coverage(cluster_masks, mask_groupings):
    """Fraction of each visibility cluster's candidate masks that ended up in a temporal group, and overall
    (cotracker_matching.py:843-872); cluster folders are zipped with the groupings positionally, as there."""
    matched_all = total_all = 0
    per_cluster = []
    for c_masks, grouping in zip(cluster_masks, mask_groupings):
        if len(c_masks) == 0:
            print("No cluster masks found for this cluster.")
            continue
        have = [(int(m["frame_id"]), int(m["mask_id"])) for m in c_masks]
        grouped = [fm for fms in grouping["overall_mask_ids_per_label"].values() for fm in fms]
        matched = sum(1 for fm in grouped if fm in have)
        cov = matched / len(have) if len(have) > 0 else 0
        print(f"Cluster coverage: {cov:.2%} ({matched}/{len(have)}) for cluster masks {c_masks[0]['vis_cluster_id']}")
        per_cluster.append(cov)
        matched_all += matched
        total_all += len(have)
    overall = matched_all / total_all if total_all > 0 else 0
    print(f"Overall coverage {overall:.2%} ({matched_all}/{total_all})")
    return overall, per_cluster


def gather_and_save_one2x_data(matches_data, mask_groupings, visibility_group_mask_path):
    """Average one-to-many flags per visibility cluster and per temporal group, written as
    cluster_{c}/one2x_data_cluster{c}.json and video_one2x_data.json (cotracker_matching.py:875-921)."""
    per_cluster = {}
    for md in matches_data:
        per_cluster.setdefault(f"cluster_{md['cluster_id']}", []).append(md["one2x"])
    video = {}
    for g in mask_groupings:
        cid = g["cluster_id"]
        od = {"avg_one2x_cluster": np.mean(per_cluster.get(f"cluster_{cid}", []))}
        for label, fms in g["overall_mask_ids_per_label"].items():
            ent = []
            for (f, m) in fms:
                e = next((md["one2x"] for md in matches_data if md["frame_id"] == f and md["mask_id"] == m), None)
                if e is not None:
                    ent.append(e)
            avg = np.sum(ent) / len(ent) if ent else 0
            od[f"group_{label}"] = {"avg_one2x": avg, "one2x_counts": len(ent), "noisy": bool(avg > 0.5)}
        with open(os.path.join(visibility_group_mask_path, f"cluster_{cid}", f"one2x_data_cluster{cid}.json"), "w") as f:
            json.dump(od, f, indent=4)
        video[f"cluster_{cid}"] = od
    with open(os.path.join(visibility_group_mask_path, "video_one2x_data.json"), "w") as f:
        json.dump(video, f, indent=4)


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
    # windowed storage: only the frames of a query's window [v0, v1] are ever voted on, so only they are kept on the
    # device (the reference holds one query's full-length tracks at a time; a dense [Nm, T, P, 2] tensor would be tens
    # of GB for long, crowded videos). Rows that are not candidates own no tile and are never read.
    ttr = max(int(rowinfo[g, 3] - rowinfo[g, 2] + 1) for g in tracks_of)
    tstart = np.zeros(nm, np.int32)
    tracks = torch.empty((nm, ttr, pmax, 2), dtype=torch.float32, device=dev)
    npts = torch.zeros(nm, dtype=torch.int32)
    for g, t in tracks_of.items():
        ts = min(int(rowinfo[g, 2]), T - ttr)
        tstart[g] = ts
        tracks[g, :, : t.shape[1]] = t[ts:ts + ttr].to(dev, torch.float32)
        npts[g] = t.shape[1]
    b = Batch([VideoInput(labels=labels_dev, tracks=tracks, npts=npts.to(dev), tstart=torch.from_numpy(tstart).to(dev))],
              stages="LD")
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
