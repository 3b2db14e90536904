"""Drop-in for keymask_ident/cotracker_occlusions.py (stage A): per (frame, mask) the tracker's
visibility flags are reduced to a per-frame visibility fraction on the GPU (K3a), the result is
written in the reference's json schema."""
from __future__ import annotations

import glob
import json
import os
import warnings

import cv2
import numpy as np
import torch

try:
    from . import _engine, crw_utils
except ImportError:
    import _engine
    import crw_utils

load_image_robust = crw_utils.load_image_robust


def load_masks(mask_folder: str):
    """(T,H,W,1) int64 label maps; None (with a warning) when the folder holds no PNG
    (cotracker_occlusions.py:22-85)."""
    if not sorted(glob.glob(os.path.join(mask_folder, "*.png"))):
        warnings.warn(f"No .png masks found in {mask_folder!r}")
        return None
    return crw_utils.load_masks(mask_folder)


def mp4_from_images(img_folder: str, frame_rate: int = 30) -> torch.Tensor:
    """(1,T,C,H,W) float video tensor from a folder of frames (cotracker_occlusions.py:88-130)."""
    paths = []
    for e in ("*.png", "*.jpg", "*.jpeg", "*.bmp"):
        paths.extend(glob.glob(os.path.join(img_folder, e)))
    paths = sorted(paths)
    if not paths:
        raise ValueError(f"No images found in {img_folder!r}")
    frames = [cv2.cvtColor(im, cv2.COLOR_BGR2RGB) for im in (cv2.imread(p) for p in paths) if im is not None]
    if not frames:
        raise ValueError("None of the images could be read successfully.")
    return torch.from_numpy(np.stack(frames, axis=0)).permute(0, 3, 1, 2)[None].float()


def get_segmentation_mask(masks: torch.Tensor, query_frame_idx: int, object_id: int = 1) -> torch.Tensor:
    """(1,1,H,W) uint8 0/255 mask of one object (or all objects for -1) of a (T,H,W,1) label tensor."""
    frame = masks[query_frame_idx, ..., 0]
    sel = (frame != 0) if object_id == -1 else (frame == object_id)
    return (sel.to(torch.uint8) * 255)[None, None]


def boolean_visibility(vis: torch.Tensor, threshold: float = 0.3) -> torch.Tensor:
    """vis >= threshold on the GPU (cotracker_occlusions.py:226-240)."""
    return _engine.boolean_visibility(vis, threshold)


def extract_appearance_events(vis: torch.Tensor, smoothing_window: int = 1, thresh: float = 0.95, min_run_length: int = 4):
    """{row: [(appear_frame, disappear_frame), ...]} of (N, T) visibility curves
    (cotracker_occlusions.py:166-223); K3d appearance_events_kernel."""
    return _engine.appearance_events(vis, smoothing_window, thresh, min_run_length)


def extract_object_visibility_data(video_path, masks_path, video_output_dir, visibility_maps_base_output_dir, debug=False):
    """Same contract as the reference (cotracker_occlusions.py:243-396): returns
    {"video_data": [{"frame_id", "data": [{"object_id", "visibility": [T floats]}]}]} or None and
    writes <base>/<dataset>/<split>/data/<video>.json."""
    masks = _engine.cached_labels(masks_path, load_masks)
    if masks is None:
        print("Failed to load masks for visibility analysis...")
        return None
    dataset_name, split = _engine.dataset_and_split(video_path)
    out_dir = os.path.join(visibility_maps_base_output_dir, dataset_name, split)
    try:
        os.makedirs(out_dir, exist_ok=True)
    except Exception as e:  # noqa: BLE001
        print(f"Failed to create visibility maps output directory {out_dir}: {e}")
        return None
    video_name = os.path.basename(video_path)
    video = mp4_from_images(video_path)
    model = _engine.make_tracker()
    if torch.cuda.is_available():
        video = video.cuda()

    labels_dev = _engine.labels_u8_device(masks)
    qframe, qlabel, _ = _engine.enumerate_objects(labels_dev)          # K0: sort(unique(label[t]))[1:]
    T = video.shape[1]
    grid_size = 50
    vis_rows = []
    for f, oid in zip(qframe.tolist(), qlabel.tolist()):
        segm = get_segmentation_mask(masks, f, object_id=oid)
        _tracks, vis = model(video, grid_size=grid_size, grid_query_frame=f, segm_mask=segm, backward_tracking=f > 0)
        vis_rows.append(vis[0])
    if not vis_rows:
        return None
    V = _engine.visibility_mean(vis_rows, T)                            # K3a on the GPU

    video_data, row = [], 0
    for f in sorted(set(qframe.tolist())):
        data = []
        while row < len(qframe) and qframe[row] == f:
            data.append({"object_id": int(qlabel[row]), "visibility": [float(x) for x in V[row]]})
            row += 1
        video_data.append({"frame_id": int(f), "data": data})
    path = os.path.join(out_dir, "data", video_name + ".json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump({"video_data": video_data}, fh, indent=4)
    return {"video_data": video_data}
