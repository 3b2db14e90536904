"""Stress sweep of the point-votes kernel (BASELINE.json configs[4] scaled to one GPU):
masks/frame 10-100, tracks 1k-16k, window 8-64 frames, windowed track storage. For each
combination: ms per launch, achieved GB/s vs the measured HBM peak, and a parity spot check against
the CPU oracle. Run on the GPU box:  python tools/sweep.py > gpurun_out/sweep.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import keymask_oracle as ko  # noqa: E402  (checker only)
from s2d_b200 import _lib  # noqa: E402
from s2d_b200.pipeline import Batch, VideoInput  # noqa: E402
from s2d_b200.synth import make_scene_device  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 6548.2
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    T, H, W = 64, 480, 854
    grid = [(M, P, Tw) for M in (10, 20, 50, 100) for P in (1024, 4096, 16384) for Tw in (8, 16, 32, 64)]
    if len(sys.argv) > 1:        # one custom point: python tools/sweep.py T H W M P Tw   (e.g. the SA-V-shaped C3 tile: 48 1080 1920 30 8192 32)
        T, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
        grid = [(int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]))]
    rows = []
    for M, P, Tw in grid:
        for _once in (0,):
            for _once2 in (0,):
                if M * T * Tw * P * 8 > 40e9:
                    continue
                sc = make_scene_device(7 + M, T, H, W, M, P, dev)
                Nm = sc["tracks"].shape[0]
                qf = sc["query_frame"].cpu().numpy()
                t0 = np.clip(qf - Tw // 2, 0, T - Tw).astype(np.int32)
                idx = torch.from_numpy(t0[:, None] + np.arange(Tw)[None, :]).to(dev).long()
                win = torch.gather(sc["tracks"], 1, idx[:, :, None, None].expand(Nm, Tw, P, 2)).contiguous()
                del sc["tracks"], sc["vis"]
                vid = VideoInput(labels=sc["labels"], tracks=win, tstart=torch.from_numpy(t0).to(dev), max_label=M)
                b = Batch([vid], stages="LD")
                ri = np.stack([np.zeros(Nm), np.zeros(Nm), t0, t0 + Tw - 1], axis=1).astype(np.int32)
                b.upload_stage_b(ri, 1, 1)
                b.run(stages="L")
                st = torch.cuda.current_stream().cuda_stream

                def votes():
                    _lib.call("s2d_point_votes", b.descs.data_ptr(), 1, b.max_T, b.max_Nm, b.max_P, b.vec4, b.total_rows,
                              b.rowinfo.data_ptr(), b.vidinfo.data_ptr(), b.pvwork.data_ptr(),
                              b.pvtmaps.data_ptr() if b.pvtmaps is not None else None, b.hits.data_ptr(),
                              b.uniq.data_ptr(), st)
                for _ in range(2):
                    votes()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    votes()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                tiles = Nm * Tw
                byts = 8.0 * P * tiles + T * H * W + 4.0 * (M + 2) * tiles
                # parity spot check: two queries, their whole window
                L = b.host_descs[0].L
                hits = b.hits.cpu().numpy().reshape(Nm, T, L)
                uniq = b.uniq.cpu().numpy().reshape(Nm, T)
                lab_h = sc["labels"].cpu().numpy()
                ok = True
                for q in (0, Nm // 2):
                    full = np.zeros((T, P, 2), np.float32)
                    full[t0[q]:t0[q] + Tw] = win[q].cpu().numpy()
                    h, u = ko.point_votes(full, lab_h, int(t0[q]), int(t0[q] + Tw - 1), nbins=L)
                    ok &= bool(np.array_equal(u, uniq[q, t0[q]:t0[q] + Tw]) and np.array_equal(h, hits[q, t0[q]:t0[q] + Tw]))
                rows.append({"masks_per_frame": M, "tracks": P, "window": Tw, "queries": int(Nm), "tiles": int(tiles),
                             "ms": ms, "GBps": byts / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": byts / (ms * 1e-3) / 1e9 / peak,
                             "frames_per_s": T / (ms * 1e-3), "parity": "ok" if ok else "MISMATCH"})
                print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
                del b, vid, win, sc
                torch.cuda.empty_cache()
    print(json.dumps({"shape": [T, H, W], "peak_gbs": peak, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
