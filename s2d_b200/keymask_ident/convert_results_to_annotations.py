"""Drop-in for keymask_ident/convert_results_to_annotations.py: turns a model's results.json (per
prediction: video_id, score, category_id, per-frame COCO RLE segmentations) into a YTVIS annotation
file, computing each frame's bounding box and area from the RLE. The reference calls
mask_util.toBbox / mask_util.area per frame (convert_results_to_annotations.py:70-81); here all RLEs
of the results file go through one GPU launch (s2d_rle_area_bbox, run-length domain, no decode).
Same CLI, same output schema (convert_results_to_annotations.py:38-44, 84-95)."""
from __future__ import annotations

import argparse
import json
import os

from tqdm import tqdm

try:
    from .annotations import rle_area_bbox_batch
except ImportError:
    from annotations import rle_area_bbox_batch


def convert_results_to_annotation(annotation_file_path, gt_annotation_path, results_file_path, score_threshold, output_dir, filename):
    """Same arguments and file written as the reference's function (which reads the score threshold from the
    script's global `args`: identical when run as a script, which is the only way the reference runs it)."""
    with open(annotation_file_path) as f:
        merged = json.load(f)
    with open(results_file_path) as f:
        results = json.load(f)
    with open(gt_annotation_path) as f:
        gt = json.load(f)
    meta = {v["id"]: v for v in gt["videos"]}
    out = {"info": gt["info"], "licenses": gt["licenses"], "videos": gt["videos"], "categories": merged["categories"],
           "annotations": []}
    low = 0
    kept, rles, where = [], [], []
    for i, pred in enumerate(tqdm(results, desc=f"Converting annotations for {os.path.basename(results_file_path)}")):
        vid = pred["video_id"]
        if pred["score"] < score_threshold:
            print(f"Skipping prediction for video {vid} due to low score: {pred['score']}")
            low += 1
            continue
        if vid not in meta:
            continue
        n = meta[vid]["length"]
        assert n == len(pred["segmentations"]), \
            f"Number of frames in video {vid} ({n}) does not match the number of segmentations ({len(pred['segmentations'])})"
        ann = {"video_id": vid, "iscrowd": 0, "height": meta[vid]["height"], "width": meta[vid]["width"], "length": n,
               "segmentations": pred["segmentations"], "bboxes": [None] * n, "areas": [None] * n,
               "category_id": pred["category_id"], "id": i + 1}
        for t, rle in enumerate(pred["segmentations"]):
            if rle is not None:
                rles.append(rle)
                where.append((len(kept), t))
        kept.append(ann)
    for (k, t), (area, box) in zip(where, rle_area_bbox_batch(rles)):
        kept[k]["bboxes"][t] = box
        kept[k]["areas"][t] = area
    out["annotations"] = kept
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, f"{filename}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=2)
    print(f"Successfully converted '{results_file_path}' to '{path}'")
    print(f"Skipped {low}/{len(results)} ({round((low / len(results)) * 100, 2)}%) low scoring predictions.")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Convert a results.json file to a COCO-style annotation file.")
    parser.add_argument("--annotation-file", help="Path to the merged.json file.")
    parser.add_argument("--gt-annotation-file", help="Root directory for the dataset, used to locate video files.")
    parser.add_argument("--results-file", help="Path to the results.json file.")
    parser.add_argument("--score-threshold", type=float, default=0.75, help="Score threshold.")
    parser.add_argument("--output-dir", help="Path to save the output annotation file.")
    parser.add_argument("--output-filename", help="File name.")
    args = parser.parse_args()
    convert_results_to_annotation(args.annotation_file, args.gt_annotation_file, args.results_file, args.score_threshold,
                                  args.output_dir, args.output_filename)
