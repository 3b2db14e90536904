"""Driver equivalent to the reference's main_keymask_ident.py (same flags, same per-stage
failure handling, same skip-if-annotation-exists resume rule, same final report). The reference's
own driver works unchanged as well when this directory precedes it on sys.path."""
from __future__ import annotations

import os

try:
    from . import _engine, crw_utils
    from .annotations import write_annotation_for_video
    from .cotracker_matching import temporal_correspondence_match
    from .cotracker_occlusions import extract_object_visibility_data
    from .identify_visibility_windows import get_visibility_windows_for_video
    from .keymask_utils import save_segmentation_masks
except ImportError:
    import _engine
    import crw_utils
    from annotations import write_annotation_for_video
    from cotracker_matching import temporal_correspondence_match
    from cotracker_occlusions import extract_object_visibility_data
    from identify_visibility_windows import get_visibility_windows_for_video
    from keymask_utils import save_segmentation_masks

# split naming of the driver differs from the stage functions' (main_keymask_ident.py:39-73)
_DRIVER_SPLIT = {"DAVIS": lambda p: "all", "VIPSeg": lambda p: "imgs", "SA-V": lambda p: "train"}


def dataset_and_split_for_driver(video_base_path):
    name, _ = _engine.dataset_and_split(video_base_path)
    split = _DRIVER_SPLIT.get(name, lambda p: "train" if "train" in p else "valid")(video_base_path)
    return name, split


def process_video(video_path, masks_path, args, dataset_name, split):
    """one iteration of the reference's per-video loop (main_keymask_ident.py:81-139); returns
    True when an annotation was written."""
    name = os.path.basename(video_path)
    stage = "visibility data extraction"
    try:
        vis = extract_object_visibility_data(video_path, masks_path, args.video_output_dir,
                                             args.visibility_maps_output_base, args.debug)
        if vis is None:
            return False
        stage = "visibility window identification"
        windows = get_visibility_windows_for_video(vis, dataset_name, split, name,
                                                   args.visibility_clusters_output_base, args.visibility_threshold,
                                                   args.debug)
        stage = "loading frames and masks"
        imgs, imgs_orig, lbls, meta = crw_utils.load_frames_and_masks(video_path, masks_path, windows, dataset_name)
        if imgs is None:
            print("Image or Mask Loading Error has occurred. Skipping video as to not crash the entire process.")
            return False
        stage = "segmentation mask saving"
        cm_path = save_segmentation_masks(imgs, imgs_orig, lbls, meta, args.save_path, args.debug)
        stage = "temporal correspondence matching"
        status = temporal_correspondence_match(video_path, masks_path, cm_path, args.visibility_maps_output_base,
                                               args.visibility_clusters_output_base, args.matching_threshold,
                                               args.debug)
    except Exception as e:  # noqa: BLE001  (the reference catches per stage and moves on)
        print(f"Error during {stage} for video {name}: {e}")
        return False
    if status > 0:
        write_annotation_for_video(video_path, cm_path, args.annotation_output_path, windows)
        return True
    print("No valid annotations found for video:", name)
    return False


def main():
    args = crw_utils.keymask_args()
    base = args.video_base_path
    names = sorted(os.listdir(base))
    videos = [os.path.join(base, n) for n in names if os.path.isdir(os.path.join(base, n))]
    masks = [os.path.join(args.mask_base_path, n) for n in names if os.path.isdir(os.path.join(args.mask_base_path, n))]
    if args.videos_per_job > 0:
        start = args.job_id * args.videos_per_job if args.job_id > 0 else 0
        videos, masks = videos[start:start + args.videos_per_job], masks[start:start + args.videos_per_job]
    dataset_name, split = dataset_and_split_for_driver(base)
    failed = 0
    for video_path, masks_path in zip(videos, masks):
        name = os.path.basename(video_path)
        if os.path.exists(os.path.join(args.annotation_output_path, f"{name}.json")):
            print(f"Annotation for video {name} already exists. Skipping.")
            continue
        if not process_video(video_path, masks_path, args, dataset_name, split):
            failed += 1
    print(f"Final Report: Successful annotations ->{len(videos) - failed}/{len(videos)}; "
          f"Failed annotations ->{failed}/{len(videos)}. ")


if __name__ == "__main__":
    main()
