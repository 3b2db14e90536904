"""Drop-in for keymask_ident/annotations.py (stage E): single-video YTVIS json from the
`cluster_*/group_*/frame{f}_mask{m}.png` tree. The reference calls pycocotools for the RLE;
pycocotools is used when importable, otherwise the same COCO RLE format is produced by the
restatement in `coco_rle` below (maskApi.c rleEncode / rleToString / rleArea / rleToBbox).
Parity at the RLE boundary is unpinned in the build container (pycocotools is not installed)."""
from __future__ import annotations

import json
import os
import re

import numpy as np
from PIL import Image


class coco_rle:
    @staticmethod
    def counts(mask: np.ndarray):
        """column-major run lengths starting with a run of zeros."""
        flat = np.asarray(mask, dtype=np.uint8).reshape(-1, order="F")
        if flat.size == 0:
            return [0]
        change = np.nonzero(flat[1:] != flat[:-1])[0] + 1
        edges = np.concatenate(([0], change, [flat.size]))
        runs = np.diff(edges).tolist()
        if flat[0] != 0:
            runs = [0] + runs
        return runs

    @staticmethod
    def to_string(cnts):
        out = []
        for i, x in enumerate(cnts):
            x = int(x)
            if i > 2:
                x -= int(cnts[i - 2])
            more = True
            while more:
                c = x & 0x1F
                x >>= 5
                more = (x != -1) if (c & 0x10) else (x != 0)
                if more:
                    c |= 0x20
                out.append(chr(c + 48))
        return "".join(out)

    @staticmethod
    def from_string(s):
        cnts, p = [], 0
        while p < len(s):
            x, k, more = 0, 0, True
            while more:
                c = ord(s[p]) - 48
                x |= (c & 0x1F) << (5 * k)
                more = bool(c & 0x20)
                p += 1
                k += 1
                if not more and (c & 0x10):
                    x |= -1 << (5 * k)
            if len(cnts) > 2:
                x += cnts[-2]
            cnts.append(x)
        return cnts

    @staticmethod
    def encode(mask: np.ndarray):
        h, w = mask.shape[:2]
        return {"size": [int(h), int(w)], "counts": coco_rle.to_string(coco_rle.counts(mask.reshape(h, w)))}

    @staticmethod
    def area(mask: np.ndarray):
        return int(np.count_nonzero(mask))

    @staticmethod
    def bbox(mask: np.ndarray):
        m = np.asarray(mask).reshape(mask.shape[0], mask.shape[1]) != 0
        if not m.any():
            return [0.0, 0.0, 0.0, 0.0]
        ys, xs = np.nonzero(m.any(axis=1))[0], np.nonzero(m.any(axis=0))[0]
        return [float(xs[0]), float(ys[0]), float(xs[-1] - xs[0] + 1), float(ys[-1] - ys[0] + 1)]


def _encode(binary_mask):
    try:
        from pycocotools import mask as mask_util
        rle = mask_util.encode(np.array(binary_mask[..., None], order="F", dtype="uint8"))[0]
        rle["counts"] = rle["counts"].decode("ascii")
        return rle, int(mask_util.area(rle)), mask_util.toBbox(rle).tolist()
    except ImportError:
        return coco_rle.encode(binary_mask), coco_rle.area(binary_mask), coco_rle.bbox(binary_mask)


def write_annotation_for_video(video_path, cluster_masks_path, annotation_output_path, visibility_data):
    """Same contract as the reference (annotations.py:8-140)."""
    video_name = os.path.basename(video_path)
    video_files = sorted(f for f in os.listdir(video_path) if f.endswith((".jpg", ".png", ".jpeg")))
    if not video_files:
        print(f"No image files found in {video_path}")
        return
    with Image.open(os.path.join(video_path, video_files[0])) as img:
        width, height = img.size
    video = {"license": 1, "coco_url": "", "height": height, "width": width, "length": len(video_files),
             "date_captured": "2019-04-11 00:55:41.903902",
             "file_names": [os.path.join(video_name, f) for f in video_files], "flickr_url": "", "id": 1}
    cluster_dirs = sorted(d for d in os.listdir(cluster_masks_path)
                          if os.path.isdir(os.path.join(cluster_masks_path, d)) and d.startswith("cluster_")
                          and any(f.endswith(".png") for f in os.listdir(os.path.join(cluster_masks_path, d))))
    with open(os.path.join(cluster_masks_path, "video_one2x_data.json")) as f:
        one2x_data = json.load(f)
    annotations, ann_id, n = [], 1, len(video_files)
    for cname in cluster_dirs:
        cdir = os.path.join(cluster_masks_path, cname)
        groups = sorted(d for d in os.listdir(cdir) if os.path.isdir(os.path.join(cdir, d)) and d.startswith("group_"))
        try:
            cid = int(cname.replace("cluster_", ""))
            ranges = next((c for c in visibility_data["clusters"] if c["cluster_id"] == cid), None)["ranges"]
        except KeyError:
            ranges = [(-1, -1)]
        if cname not in one2x_data:
            print(f"Could not find one2x data for {cname}.")
            continue
        for gname in groups:
            gdir = os.path.join(cdir, gname)
            segs, boxes, areas = [None] * n, [None] * n, [None] * n
            for mf in (f for f in os.listdir(gdir) if f.endswith(".png")):
                m = re.search(r"frame(\d+)", mf)
                if not m or int(m.group(1)) >= n:
                    continue
                binary = np.array(Image.open(os.path.join(gdir, mf)).convert("L")) > 0
                segs[int(m.group(1))], areas[int(m.group(1))], boxes[int(m.group(1))] = _encode(binary)
            annotations.append({"video_id": 1, "iscrowd": 0, "height": height, "width": width, "length": n,
                                "segmentations": segs, "bboxes": boxes, "areas": areas, "category_id": 1,
                                "id": ann_id, "one2x": round(float(one2x_data[cname][gname]["avg_one2x"]), 2),
                                "visibility_ranges": ranges})
            ann_id += 1
    os.makedirs(annotation_output_path, exist_ok=True)
    with open(os.path.join(annotation_output_path, f"{video_name}.json"), "w") as f:
        json.dump({"videos": [video], "annotations": annotations,
                   "categories": [{"supercategory": "object", "id": 1, "name": "fg"}]}, f)
