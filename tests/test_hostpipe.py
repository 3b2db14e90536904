"""End-to-end entry for host buffers (s2d_b200/hostpipe.py, hostmem.py): same result tables as the device-resident
batch, whatever the wire format of the visibility flags and the kind of pinned pool."""
import numpy as np
import pytest
import torch

from s2d_b200 import hostmem


def test_cpulist_parser_and_binding_report():
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert hostmem._parse_cpulist("") == []


@pytest.mark.gpu
@pytest.mark.parametrize("bits,huge", [(False, False), (True, True)])
def test_host_pipeline_matches_device_batch(bits, huge):
    from s2d_b200.hostpipe import HostPipeline, HostVideo
    from s2d_b200.pipeline import Batch, VideoInput, pack_vis_bits
    from s2d_b200.synth import make_scene
    dev = torch.device("cuda:0")
    shapes = [(12, 64, 96, 4, 64), (10, 48, 80, 3, 32), (12, 64, 96, 4, 64), (10, 48, 80, 3, 32)]   # slots 0 / 1 of two chunks
    # the two videos that land in a staging slot must have the same shapes (same number of queries): chunk 1 reuses
    # chunk 0's scenes with other tracks
    scenes = [make_scene(300 + (i % 2), T, H, W, M, P) for i, (T, H, W, M, P) in enumerate(shapes)]
    scenes[2].tracks[:] = scenes[0].tracks[::-1].copy()
    scenes[3].tracks[:] = scenes[1].tracks + 0.25
    want = []
    for sc in scenes:
        b = Batch([VideoInput(torch.from_numpy(sc.labels).to(dev), torch.from_numpy(sc.tracks).to(dev), torch.from_numpy(sc.vis).to(dev))])
        b.run()
        torch.cuda.synchronize()
        want.append(b.fetch_summary())
    tensors = []
    for sc in scenes:
        v = torch.from_numpy(sc.vis)
        tensors += [torch.from_numpy(sc.labels), torch.from_numpy(sc.tracks), pack_vis_bits(v) if bits else v]
    if huge:
        pool, host = hostmem.pooled_copies(tensors, huge=True)
        assert pool.registered and all(t.is_pinned() for t in host)
    else:
        pool, host = None, [t.pin_memory() for t in tensors]
    hv = [HostVideo(host[3 * i], host[3 * i + 1], host[3 * i + 2], vis_bits=bits) for i in range(4)]
    pipe = HostPipeline(hv[:2], dev, chunk=2)
    for _ in range(2):                                   # second run reuses the staging sets and result tables
        outs = pipe.run(hv)
    assert len(outs) == 2 and pipe.h2d_bytes == sum(v.nbytes for v in hv) and pipe.d2h_bytes > 0
    for c in range(2):
        for j in range(2):
            ref = want[2 * c + j]
            nm = ref["rowinfo"].shape[0]
            row0 = 0 if j == 0 else want[2 * c]["rowinfo"].shape[0]
            got = outs[c]
            assert np.array_equal(got["vidinfo"].numpy().reshape(2, -1)[j], ref["vidinfo"][0])
            assert np.array_equal(got["rowinfo"].numpy().reshape(-1, 4)[row0:row0 + nm], ref["rowinfo"])
            assert np.array_equal(got["glabel"].numpy()[row0:row0 + nm], ref["glabel"])
            assert np.array_equal(got["one2x"].numpy()[row0:row0 + nm], ref["one2x"])
    if pool is not None:
        pool.close()
