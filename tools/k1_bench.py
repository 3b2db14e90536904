"""K1 on the tensor cores: achieved int8 TOP/s of the one-hot Gram kernel vs a measured int8 GEMM
peak (torch._int_mm 8192^3, cuBLASLt) and the nominal 4.5 POP/s. Run on the GPU box."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2d_b200 import _lib  # noqa: E402


def ev_time(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device("cuda:0")
    out = {}
    n = 8192
    x = torch.randint(-3, 3, (n, n), dtype=torch.int8, device=dev)
    y = torch.randint(-3, 3, (n, n), dtype=torch.int8, device=dev)
    ms = ev_time(lambda: torch._int_mm(x, y))
    out["int8_cublaslt_8192_tops"] = 2 * n ** 3 / (ms * 1e-3) / 1e12
    del x, y
    st = torch.cuda.current_stream().cuda_stream
    from s2d_b200.synth import make_scene_device
    for (F, L, H, W, nv, kind) in [(36, 21, 720, 1280, 8, "scene"), (36, 21, 480, 854, 8, "scene"),
                                   (64, 31, 1080, 1920, 2, "scene"), (36, 21, 720, 1280, 8, "random")]:
        if kind == "random":      # worst case: a different label at every pixel
            labs = [torch.randint(0, L, (F, H, W), dtype=torch.uint8, device=dev) for _ in range(nv)]
        else:                     # piecewise-constant label maps of the synthetic scenes (P=2: tracks are not used here)
            labs = [make_scene_device(2024 + i, F, H, W, L - 1, 2, dev)["labels"] for i in range(nv)]
        R = F * L
        G = torch.empty(R * R, dtype=torch.int32, device=dev)
        import ctypes as C
        nw = C.c_int64()
        _lib.call("s2d_overlap_gram_work_ints", F, L, H * W, C.byref(nw))
        work = torch.empty(nw.value, dtype=torch.int32, device=dev)

        def run():
            for lab in labs:
                _lib.call("s2d_overlap_gram_labels", lab.data_ptr(), F, L, H * W, work.data_ptr(), G.data_ptr(), st)
        ms = ev_time(run, iters=5, warm=2) / nv
        ops = 2.0 * R * R * H * W
        out[f"gram_F{F}_L{L}_{H}x{W}_{kind}"] = {"ms_per_video": ms, "rows": R, "tops": ops / (ms * 1e-3) / 1e12,
                                          "label_bytes": F * H * W, "hbm_gbs": F * H * W / (ms * 1e-3) / 1e9}
    peak = out["int8_cublaslt_8192_tops"]
    for k, v in out.items():
        if isinstance(v, dict):
            v["frac_of_measured_int8_peak"] = v["tops"] / peak
            v["frac_of_nominal_4500"] = v["tops"] / 4500.0
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
