"""Drop-in for keymask_ident/annotations.py (stage E): single-video YTVIS json from the
`cluster_*/group_*/frame{f}_mask{m}.png` tree. The reference calls pycocotools per keymask
(mask_util.encode / area / toBbox, annotations.py:94-106); here all keymasks of a video are encoded
in one batch on the GPU (s2d_rle_encode: column-major run lengths, area, bounding box) and only the
base-48 string packing of the short count lists (maskApi.c rleToString) runs on the host.
pycocotools is not installable in the build container, so parity at the RLE boundary is pinned by
the published algorithm and by decode round trips only (tests/test_dropin_gpu.py)."""
from __future__ import annotations

import ctypes as C
import json
import os
import re

import numpy as np
import torch
from PIL import Image

try:
    from . import _engine
except ImportError:
    import _engine


class coco_rle:
    """COCO compressed-RLE strings (maskApi.c rleToString / rleFrString) and the GPU batch encoder."""

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
    def encode_batch(masks):
        """masks [N,H,W] (bool / uint8, numpy or torch) -> list of (rle dict, area int, bbox [x,y,w,h] floats),
        what mask_util.encode / area / toBbox return per mask."""
        from s2d_b200 import _lib
        dev = _engine.device()
        m = torch.as_tensor(np.asarray(masks)) if not torch.is_tensor(masks) else masks
        m = (m != 0).to(dev, torch.uint8).contiguous()
        n, h, w = m.shape
        if n == 0:
            return []
        st = torch.cuda.current_stream(dev).cuda_stream
        max_runs = min(h * w + 1, 8 * w + 64)
        while True:
            need = C.c_int64()
            _lib.call("s2d_rle_work_ints", n, h, w, max_runs, C.byref(need))
            work = torch.empty(need.value, dtype=torch.int32, device=dev)
            counts = torch.empty((n, max_runs), dtype=torch.int32, device=dev)
            nruns = torch.empty(n, dtype=torch.int32, device=dev)
            area = torch.empty(n, dtype=torch.int32, device=dev)
            bbox = torch.empty((n, 4), dtype=torch.int32, device=dev)
            _lib.call("s2d_rle_encode", m.data_ptr(), n, h, w, max_runs, work.data_ptr(), counts.data_ptr(),
                      nruns.data_ptr(), area.data_ptr(), bbox.data_ptr(), st)
            nr = nruns.cpu().numpy()
            if int(nr.max()) <= max_runs:
                break
            max_runs = int(nr.max())                   # a mask with more runs than the first guess: once more
        cn, ar, bb = counts.cpu().numpy(), area.cpu().numpy(), bbox.cpu().numpy()
        out = []
        for i in range(n):
            rle = {"size": [int(h), int(w)], "counts": coco_rle.to_string(cn[i, :nr[i]].tolist())}
            if ar[i] == 0:
                box = [0.0, 0.0, 0.0, 0.0]
            else:
                box = [float(bb[i, 0]), float(bb[i, 1]), float(bb[i, 2] - bb[i, 0] + 1), float(bb[i, 3] - bb[i, 1] + 1)]
            out.append((rle, int(ar[i]), box))
        return out


def rle_area_bbox_batch(rles):
    """[(area int, bbox [x, y, w, h] floats)] of COCO RLE dicts ({"size": [h, w], "counts": str | bytes | list})
    without decoding the masks: mask_util.area / toBbox for a whole list in one GPU launch (s2d_rle_area_bbox)."""
    from s2d_b200 import _lib
    if not rles:
        return []
    dev = _engine.device()
    lists = []
    for r in rles:
        c = r["counts"]
        if isinstance(c, bytes):
            c = c.decode("ascii")
        lists.append(coco_rle.from_string(c) if isinstance(c, str) else [int(x) for x in c])
    offs = np.zeros(len(lists) + 1, np.int64)
    offs[1:] = np.cumsum([len(x) for x in lists])
    flat = np.concatenate([np.asarray(x, np.int64) for x in lists] + [np.zeros(1, np.int64)]).astype(np.int32)
    counts = torch.from_numpy(flat).to(dev)
    offsets = torch.from_numpy(offs).to(dev)
    hs = torch.tensor([int(r["size"][0]) for r in rles], dtype=torch.int32, device=dev)
    area = torch.empty(len(rles), dtype=torch.int32, device=dev)
    bbox = torch.empty((len(rles), 4), dtype=torch.int32, device=dev)
    _lib.call("s2d_rle_area_bbox", counts.data_ptr(), offsets.data_ptr(), len(rles), hs.data_ptr(), area.data_ptr(),
              bbox.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    a, b = area.cpu().numpy(), bbox.cpu().numpy()
    return [(int(a[i]), [float(v) for v in b[i]]) for i in range(len(rles))]


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
    pending = []
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
                pending.append((segs, areas, boxes, int(m.group(1)), binary))     # encoded in one GPU batch below
            annotations.append({"video_id": 1, "iscrowd": 0, "height": height, "width": width, "length": n,
                                "segmentations": segs, "bboxes": boxes, "areas": areas, "category_id": 1,
                                "id": ann_id, "one2x": round(float(one2x_data[cname][gname]["avg_one2x"]), 2),
                                "visibility_ranges": ranges})
            ann_id += 1
    # later files of a group overwrite earlier ones on the same frame, like the reference's loop
    by_shape = {}
    for item in pending:
        by_shape.setdefault(item[4].shape, []).append(item)
    for items in by_shape.values():
        enc = coco_rle.encode_batch(np.stack([it[4] for it in items]))
        for (segs, areas, boxes, fidx, _), (rle, area, box) in zip(items, enc):
            segs[fidx], areas[fidx], boxes[fidx] = rle, area, box
    os.makedirs(annotation_output_path, exist_ok=True)
    with open(os.path.join(annotation_output_path, f"{video_name}.json"), "w") as f:
        json.dump({"videos": [video], "annotations": annotations,
                   "categories": [{"supercategory": "object", "id": 1, "name": "fg"}]}, f)
