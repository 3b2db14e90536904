"""TEST INFRASTRUCTURE ONLY. Interleaved A/B timing of the reference's own extract_mask_matches
(/root/reference/keymask_ident/cotracker_matching.py:665-719, unmodified, imported through the stubs of
oracle/ref_harness.py) against the dense torch-CPU port bench.py runs on the GPU box (oracle/dense_port.py), on the
same queries of a C1-shaped scene. The build container's cores are shared, so whole-run wall times move by tens of
percent between runs: here the two implementations alternate query by query and the minimum over ROUNDS rounds counts.

    python oracle/time_port_vs_reference.py [nqueries] [out.json]          (build container only)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import dense_port, ref_harness  # noqa: E402
from s2d_b200.synth import make_scene  # noqa: E402


ROUNDS = 7


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    T, H, W, M, P = 24, 480, 854, 10, 1000
    scene = make_scene(1234, T, H, W, M, P)
    ref_harness._install_stubs(lambda checkpoint=None: None)
    if ref_harness.REFERENCE_DIR not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_DIR)
    import cotracker_matching as cm                                   # the unmodified reference module

    masks = torch.from_numpy(scene.labels.astype(np.int64))[..., None]      # (T,H,W,1) int64, as load_masks returns
    glob_lut = cm.contruct_frameid_maskid_lookup(masks)
    rows = np.linspace(0, len(scene.query_frame) - 1, nq).astype(int)
    clus_lut = [[{"cluster_mask_id": i, "frame_id": int(scene.query_frame[r]), "mask_id": int(scene.query_label[r])}
                 for i, r in enumerate(range(len(scene.query_frame)))]]
    v_range = (0, T - 1)
    best_ref, best_port = [float("inf")] * nq, [float("inf")] * nq
    same = True
    for rnd in range(ROUNDS):
        for i, r in enumerate(rows):
            f, l = int(scene.query_frame[r]), int(scene.query_label[r])
            segm = (masks[f, ..., 0] == l).to(torch.uint8) * 255
            tracks = torch.from_numpy(scene.tracks[r])[None]                  # (1,T,P,2)
            t0 = time.perf_counter()
            m_ref, c_ref = cm.extract_mask_matches(segm, tracks, masks, f, v_range, 50, glob_lut, clus_lut, 0, 0.5)
            t1 = time.perf_counter()
            m_port, c_port, _ = dense_port.match_query_dense(masks, tracks[0], v_range[0], v_range[1], H, W, 0.5)
            t2 = time.perf_counter()
            best_ref[i] = min(best_ref[i], t1 - t0)
            best_port[i] = min(best_port[i], t2 - t1)
            same &= [(c["frame_id"], c["mask_id"], c["iou"]) for c in c_ref] == [(c[0], c[1], c[4]) for c in c_port]
            same &= [(m["frame_id"], m["mask_id"]) for m in m_ref] == list(m_port)
    out = {"config": "C1-shaped scene (24 x 480 x 854, 10 masks/frame, 1000 tracks), full-video window",
           "queries": int(nq), "pairs_per_query": len(c_ref), "rounds": ROUNDS, "statistic": "sum over queries of the per-query minimum",
           "host": {"cores": os.cpu_count(), "torch_threads": torch.get_num_threads()},
           "reference_extract_mask_matches_s": sum(best_ref), "port_match_query_dense_s": sum(best_port),
           "port_over_reference_time": sum(best_port) / sum(best_ref), "same_results": bool(same)}
    text = json.dumps(out, indent=1)
    print(text)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as fh:
            fh.write(text + "\n")


if __name__ == "__main__":
    main()
