"""Drop-in for keymask_ident/identify_visibility_windows.py (stage B): binarise, Hamming DBSCAN,
majority vote, run-length windows, highly-visible rows and candidates run on the GPU (K3b/K3c);
this module keeps the json schema and file location."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

try:
    from . import _engine
except ImportError:
    import _engine


def _rows(json_file_or_data):
    if isinstance(json_file_or_data, str):
        with open(json_file_or_data) as f:
            data = json.load(f)
    else:
        data = json_file_or_data
    for fr in data["video_data"]:
        for obj in fr["data"]:
            yield fr["frame_id"], obj["object_id"], obj["visibility"]


def json_to_tensor(json_file_or_data):
    return torch.stack([torch.tensor(v) for _, _, v in _rows(json_file_or_data)])


def json_to_lookup_dict(json_file_or_data):
    return [{"frame_id": f, "object_id": o} for f, o, _ in _rows(json_file_or_data)]


def get_visible_ranges(maj_vote):
    """inclusive (start, end) runs of ones - computed by the K3c kernel."""
    m = np.asarray(maj_vote.cpu() if torch.is_tensor(maj_vote) else maj_vote).astype(bool)
    V = np.where(m[None, :], 1.0, 0.0).astype(np.float32)
    V = np.repeat(V, 5, axis=0)                      # five identical rows -> one cluster whose majority is m
    cl = _engine.visibility_windows(V, np.zeros(5, np.int32), np.arange(5, dtype=np.int32), 0.5)
    return [tuple(r) for r in cl[0]["ranges"]] if cl else []


def get_highly_visible_rows(cluster_vis, runs, threshold=0.8):
    out = {}
    cv = cluster_vis if torch.is_tensor(cluster_vis) else torch.as_tensor(cluster_vis)
    for (s, e) in runs:
        frac = cv[:, s:e + 1].sum(dim=1) / (e - s + 1)
        out[(s, e)] = (frac > threshold).nonzero(as_tuple=True)[0].tolist()
    return out


def boolean_visibility(vis: torch.Tensor, threshold: float = 0.3) -> torch.Tensor:
    return vis >= threshold


def get_visibility_windows_for_video(video_data, dataset_name, split, video_name, cluster_output_dir,
                                     visibility_threshold, debug=False):
    """Same contract as the reference (identify_visibility_windows.py:108-231)."""
    rows = list(_rows(video_data))
    V = np.asarray([v for _, _, v in rows], dtype=np.float32)
    qframe = np.asarray([f for f, _, _ in rows], np.int32)
    qlabel = np.asarray([o for _, o, _ in rows], np.int32)
    clusters = _engine.visibility_windows(V, qframe, qlabel, visibility_threshold)
    for c in clusters:      # the reference stores tuples, which json turns into lists
        c["ranges"] = [tuple(r) for r in c["ranges"]]
        for ac in c["all_candidates"]:
            ac["range"] = tuple(ac["range"])
    out = {"video_name": video_name, "clusters": clusters}
    path = f"{cluster_output_dir}/{dataset_name}/{split}/{video_name}.json"
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=4)
    print("Saved visibility clusters for video:", video_name, "to", path)
    return out
