"""Drop-in for keymask_ident/keymask_utils.py (stage C): candidate masks -> 0/255 PNGs."""
from __future__ import annotations

import os

import torch
from PIL import Image


def extract_visibility_data(visibility_data):
    clusters = [[{"range": c["range"], "mask_candidates": c["candidates"]} for c in cl["all_candidates"]]
                for cl in visibility_data["clusters"]]
    return clusters, visibility_data["video_name"]


def get_segmentation_mask(masks: torch.Tensor, query_frame_idx: int, object_id: int = 1) -> torch.Tensor:
    """(1,1,H,W) uint8 0/255 mask from a (T,1,H,W) label tensor (keymask_utils.py:37-67)."""
    frame = masks[query_frame_idx, 0]
    sel = (frame != 0) if object_id == -1 else (frame == object_id)
    return (sel.to(torch.uint8) * 255)[None, None]


def save_segmentation_masks(imgs, imgs_orig, lbls, meta, save_dir, debug=False):
    """writes <save_dir>/<video>/cluster_{c}/cluster{c}_frame{f}_mask{m}.png for every candidate
    and returns <save_dir>/<video> (keymask_utils.py:70-128)."""
    clusters, video_name = extract_visibility_data(meta["visibility"])
    out = os.path.join(save_dir, video_name)
    os.makedirs(out, exist_ok=True)
    for cid, ranges in enumerate(clusters):
        for rng in ranges:
            for cand in rng["mask_candidates"]:
                f, m = cand["frame_id"], cand["mask_id"]
                mask = get_segmentation_mask(lbls, f, object_id=m).squeeze().cpu().numpy()
                cdir = os.path.join(out, f"cluster_{cid}")
                os.makedirs(cdir, exist_ok=True)
                Image.fromarray(mask).save(os.path.join(cdir, f"cluster{cid}_frame{f}_mask{m}.png"))
    return out
