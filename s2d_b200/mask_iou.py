"""Second consumer of K1: the dense mask x mask IoU / IoY matrices that the
reference's training-side copy-paste augmentation and self-training annotation merge compute with a
float matmul over pixels. Same function names, argument meaning and result dtype as the reference:

  mask_iou_matrix(x, y, mode)   model_training/mask2former_video/engine/train_loop.py:378-388
                                (identical copy: model_training/cutler/engine/train_loop.py:91-101)
  BatchIoU(masks1, masks2)      model_training/cutler/tools/get_self_training_ann.py:80-89

The intersection counts come from the bit-packed AND+popc kernel (s2d_overlap_bits) or, for large
operands, the tcgen05 kind::i8 kernel (s2d_overlap_i8); counts below 2^24 are exactly what the
reference's float32 matmul produces, and the final divisions are done in float32 like the reference's.
No CPU fallback: the CUDA library is required."""
from __future__ import annotations

import torch

from . import _lib


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("s2d_b200 needs a CUDA device: the overlap kernels have no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def overlap_counts(a: torch.Tensor, b: torch.Tensor, tensor_cores: bool | None = None):
    """a [Na, ...], b [Nb, ...] binary masks (non-zero = set) -> (I int32 [Na, Nb], |a| int32 [Na], |b| int32 [Nb])
    on the GPU."""
    dev = _device()
    st = torch.cuda.current_stream(dev).cuda_stream
    A = (a.reshape(a.shape[0], -1) != 0).to(dev, torch.uint8).contiguous()
    B = (b.reshape(b.shape[0], -1) != 0).to(dev, torch.uint8).contiguous()
    Na, npix = A.shape
    Nb = B.shape[0]
    assert B.shape[1] == npix, "masks must have the same number of pixels"
    I = torch.zeros((Na, Nb), dtype=torch.int32, device=dev)
    if Na == 0 or Nb == 0 or npix == 0:
        return I, torch.zeros(Na, dtype=torch.int32, device=dev), torch.zeros(Nb, dtype=torch.int32, device=dev)
    if tensor_cores is None:
        tensor_cores = Na * Nb >= 128 * 64 and npix % 16 == 0
    nw = (npix + 31) // 32
    ba = torch.empty((Na, nw), dtype=torch.int32, device=dev)
    bb = torch.empty((Nb, nw), dtype=torch.int32, device=dev)
    areaA = torch.empty(Na, dtype=torch.int32, device=dev)
    areaB = torch.empty(Nb, dtype=torch.int32, device=dev)
    _lib.call("s2d_pack_bits", A.data_ptr(), Na, npix, ba.data_ptr(), st)
    _lib.call("s2d_pack_bits", B.data_ptr(), Nb, npix, bb.data_ptr(), st)
    if tensor_cores and npix % 16 == 0:
        _lib.call("s2d_overlap_i8", A.data_ptr(), Na, B.data_ptr(), Nb, npix, I.data_ptr(), st)
        # the areas still come from the packed rows (one-column overlap launches)
        scratch = torch.empty(max(Na, Nb), dtype=torch.int32, device=dev)
        one = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.call("s2d_overlap_bits", ba.data_ptr(), Na, bb.data_ptr(), 1, nw, scratch.data_ptr(), areaA.data_ptr(), one.data_ptr(), st)
        _lib.call("s2d_overlap_bits", bb.data_ptr(), Nb, ba.data_ptr(), 1, nw, scratch.data_ptr(), areaB.data_ptr(), one.data_ptr(), st)
    else:
        _lib.call("s2d_overlap_bits", ba.data_ptr(), Na, bb.data_ptr(), Nb, nw, I.data_ptr(), areaA.data_ptr(),
                  areaB.data_ptr(), st)
    return I, areaA, areaB


def mask_iou_matrix(x: torch.Tensor, y: torch.Tensor, mode: str = "iou") -> torch.Tensor:
    """n1 x n2 IoU (or IoY = intersection / |y|) of binary masks x [n1,H,W], y [n2,H,W]; float32, on the
    caller's device, nan where the denominator is 0 - like the reference's x @ y.T formulation."""
    I, sx, sy = overlap_counts(x, y)
    inter = I.to(torch.float32)
    sum_x = sx.to(torch.float32)[:, None].expand(I.shape)
    sum_y = sy.to(torch.float32)[None, :].expand(I.shape)
    out = inter / sum_y if mode == "ioy" else inter / (sum_x + sum_y - inter)
    return out.to(x.device)


def BatchIoU(masks1: torch.Tensor, masks2: torch.Tensor) -> torch.Tensor:
    """n1 x n2 IoU of masks thresholded at 0.5 (get_self_training_ann.py:80-89): float32(|a and b|) / |a or b|."""
    I, sa, sb = overlap_counts(masks1 > 0.5, masks2 > 0.5)
    union = sa[:, None].to(torch.int64) + sb[None, :].to(torch.int64) - I.to(torch.int64)
    return (I.to(torch.float32) / union).to(masks1.device)
