import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2d_b200 import _lib
from s2d_b200.synth import make_scene_device
dev = torch.device("cuda:0")
F, L, H, W = 36, 21, 720, 1280
lab = make_scene_device(2024, F, H, W, L - 1, 2, dev)["labels"]
R = F * L
G = torch.empty(R * R, dtype=torch.int32, device=dev)
n = C.c_int64(); _lib.call("s2d_overlap_gram_work_ints", F, L, H * W, C.byref(n))
work = torch.empty(n.value, dtype=torch.int32, device=dev)
for _ in range(3):
    _lib.call("s2d_overlap_gram_labels", lab.data_ptr(), F, L, H * W, work.data_ptr(), G.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", int(G.sum().item()))
