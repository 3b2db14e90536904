"""A/B of the votes kernel's sparse-tile mode (block-summary label maps, P <= 1024) in ONE process on one GPU:
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_exp.so python tools/r02_bm_ab.py > gpurun_out/r02_bm_ab.json
Per workload (C1 video, a 32-video slice of the C4 mixture, three 1 k-track points of the C5 sweep) and per kernel
configuration: K2 time, fraction of the measured HBM copy peak, frames/s, and the digest of the results (must not depend
on the configuration). Configurations: the table path (block maps off), block maps at 6 CTAs x 128 threads per SM (the
product setting), and - experiments build only (S2D_PV_SMALL) - more CTAs per SM / other CTA shapes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from s2d_b200 import workloads as wl                                # noqa: E402
from s2d_b200.pipeline import Params                                # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6548.2) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6548.2


def main():
    dev = torch.device("cuda:0")
    exp = "exp" in os.environ.get("S2D_B200_LIB", "")
    lists = {
        "c1": [wl.VideoSpec("c1", 2024, 24, 480, 854, 10, 1000)],
        "c4_32": wl.c4_specs(512)[:32],
        "c5_m10_t8": wl.c5_specs(10, 1024, 8),
        "c5_m50_t32": wl.c5_specs(50, 1024, 32),
        "c5_m100_t64": wl.c5_specs(100, 1024, 64),
    }
    configs = [("table", False, None), ("bm_6x128", True, None)]
    if exp:
        configs += [(f"bm_small{v}", True, str(v)) for v in (7, 8, 1608, 1610, 1612, 408)] + [("table_small8", False, "8")]
    out = {"peak_gbs": PEAK, "lib": os.environ.get("S2D_B200_LIB", "product"), "rows": []}
    # warm up the process (allocator, module load)
    wl.DeviceRunner(dev, Params()).run_list([wl.VideoSpec(f"warm{i}", 7 + i, 16, 240, 426, 8, 1024) for i in range(2)])
    for lname, specs in lists.items():
        ref = None
        for cname, bm, small in configs:
            if small is None:
                os.environ.pop("S2D_PV_SMALL", None)
            else:
                os.environ["S2D_PV_SMALL"] = small
            runner = wl.DeviceRunner(dev, Params(), block_maps=bm)
            try:
                res, st, _ = runner.run_list(specs, reps=3)
            except Exception as e:                                  # a configuration that cannot launch is a data point too
                out["rows"].append({"list": lname, "config": cname, "error": str(e)[:200]})
                continue
            dig = wl.list_digest(res)
            ref = ref or dig
            k2ms = st["k2_ms"]                                       # of the last repetition
            row = {"list": lname, "config": cname, "k2_ms": round(k2ms, 4), "k2_frac": round(st["k2_bytes"] / (k2ms / 1e3) / 1e9 / PEAK, 4),
                   "device_ms_last_rep": round(st["device_ms"], 3), "frames_per_s": round(st["frames"] / (st["device_ms"] / 1e3)),
                   "tiles": st["k2_tiles"], "same_results": dig == ref}
            out["rows"].append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
