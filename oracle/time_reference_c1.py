"""TEST INFRASTRUCTURE ONLY. Times the *unmodified* reference (/root/reference/keymask_ident, through
oracle/ref_harness.py) on the C1 configuration of BASELINE.json - one synthetic 24-frame 480x854 video, 10 masks per
frame, 1000 tracks per query - stage by stage, with the time spent inside extract_mask_matches (the scope bench.py's CPU arm times with the port)
singled out. The container's cores are shared: walls move by tens of percent between runs; the port-versus-reference
ratio is therefore measured separately, interleaved and min-of-N, by oracle/time_port_vs_reference.py.

    python oracle/time_reference_c1.py [out.json]      (build container only; about three minutes on 8 cores)
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from s2d_b200.synth import make_scene  # noqa: E402


def main():
    T, H, W, M, P = 24, 480, 854, 10, 1000
    scene = make_scene(1234, T, H, W, M, P)
    with tempfile.TemporaryDirectory() as wd:
        t0 = time.perf_counter()
        ref = ref_harness.run_reference(scene, wd, log_pairs=False)        # timing run: nothing added to the inner loop
        wall = time.perf_counter() - t0
    tm = ref["timing"]
    queries = ref["queries"]                                   # one entry per extract_mask_matches call
    npairs = sum(len(q["comps"]) for q in queries)
    out = {
        "config": "C1: single synthetic 24-frame 480x854 video, 10 masks/frame, 1k tracks (BASELINE.json configs[0])",
        "host": {"cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "where": "build container (not the GPU box)"},
        "reference": {
            "status": ref["status"], "stage_calls": len(queries), "mask_pairs": npairs,
            "wall_s_all_stages_incl_io": wall,
            "stage_a_s": tm.get("stage_a_s"), "stage_b_s": tm.get("stage_b_s"), "stage_c_s": tm.get("stage_c_s"),
            "stage_d_s": tm.get("stage_d_s"), "extract_mask_matches_s": tm["extract_mask_matches_s"],
            "frames_per_s_extract_mask_matches_scope": T / tm["extract_mask_matches_s"],
            "note": "unmodified /root/reference/keymask_ident with a replaying fake tracker; stage walls include the "
                    "reference's PNG/JSON I/O and its three load_masks passes; extract_mask_matches_s is the scope "
                    "bench.py's CPU arm times",
        },
    }
    text = json.dumps(out, indent=1)
    print(text)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
