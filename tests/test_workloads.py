"""BASELINE.json configs 3-5 as list workloads (s2d_b200/workloads.py): host-side logic on CPU, the device runner and
the down-scaled SA-V-shaped parity case on the GPU."""
import numpy as np
import pytest
import torch

from s2d_b200 import workloads as wl


def test_spec_lists_are_deterministic_and_shaped_like_the_configs():
    a, b = wl.c4_specs(), wl.c4_specs()
    assert a == b and len(a) == 512
    assert {(s.H, s.W) for s in a} == {(480, 854), (720, 1280), (1080, 1920)}
    assert min(s.T for s in a) >= 20 and max(s.T for s in a) <= 120 and min(s.M for s in a) >= 5 and max(s.M for s in a) <= 30
    assert len({s.seed for s in a}) == 512
    c3 = wl.c3_specs()
    assert len(c3) == 16 and all((s.T, s.H, s.W, s.M, s.P, s.window, s.vis_bits) == (300, 1080, 1920, 30, 8192, 64, True) for s in c3)
    pts = wl.c5_points()
    assert len(pts) == 48 and {p[0] for p in pts} == {10, 20, 50, 100} and {p[1] for p in pts} == {1024, 4096, 16384} and {p[2] for p in pts} == {8, 16, 32, 64}
    for M, P, Tw in pts:
        sp = wl.c5_specs(M, P, Tw)
        assert 1 <= len(sp) <= 16 and all((s.T, s.H, s.W, s.M, s.P) == (Tw, 480, 854, M, P) for s in sp)


def test_chunks_respect_the_hbm_budget_and_keep_order():
    specs = wl.c4_specs(64)
    r = wl.DeviceRunner("cpu", budget_bytes=20e9)
    ch = r.chunks(specs)
    assert [i for c in ch for i in c] == list(range(64)) and len(ch) > 1
    for c in ch:
        assert len(c) == 1 or sum(specs[i].device_bytes() for i in c) <= 20e9
    assert wl.DeviceRunner("cpu", budget_bytes=1e15).chunks(specs) == [list(range(64))]


def test_window_starts_are_centred_clipped_and_inside_the_video():
    T, w = 300, 64
    qf = torch.tensor([0, 10, 150, 290, 299, 100, 100], dtype=torch.int32)
    ri = torch.tensor([[0, 0, 0, 299], [0, 0, 0, 299], [0, 0, 0, 299], [0, 0, 0, 299], [0, 0, 0, 299],
                       [0, 0, 90, 120], [0, -1, -1, -1]], dtype=torch.int32)
    ts = wl.window_starts(qf, ri, T, w)
    assert ts.tolist() == [0, 0, 118, 236, 236, 90, 0]
    assert ((ts >= 0) & (ts + w <= T)).all()
    # a window shorter than the video never starts before v0 when the cluster window is longer than it
    assert wl.window_starts(torch.tensor([5]), torch.tensor([[0, 0, 3, 200]]), T, w).tolist() == [3]


def test_digest_depends_on_every_result_table():
    summ = dict(vidinfo=np.arange(16, dtype=np.int32).reshape(2, 8), clusterinfo=np.zeros((2, 16, 16), np.int32),
                rowinfo=np.arange(40, dtype=np.int32).reshape(10, 4), glabel=np.arange(10, dtype=np.int32), one2x=np.zeros(10, np.int32))
    d0 = wl.video_digest(summ, 0, 0, 5)
    assert d0 == wl.video_digest(summ, 0, 0, 5) and d0 != wl.video_digest(summ, 1, 5, 5)
    for k in summ:
        t = {kk: v.copy() for kk, v in summ.items()}
        t[k].reshape(-1)[0] += 1
        assert wl.video_digest(t, 0, 0, 5) != d0, k
    r = [{"name": "a", "digest": d0}, {"name": "b", "digest": d0}]
    assert wl.list_digest(r) != wl.list_digest(r[::-1])


@pytest.mark.gpu
def test_device_runner_digests_do_not_depend_on_chunking():
    """the same list in one chunk and in many (tiny HBM budget): identical per-video digests, i.e. a video's result does
    not depend on what shares its batch - which is what makes the result independent of the GPU count."""
    dev = torch.device("cuda:0")
    specs = [wl.VideoSpec(f"v{i}", 900 + i, T, H, W, M, P) for i, (T, H, W, M, P) in enumerate(
        [(12, 96, 128, 4, 64), (40, 60, 80, 3, 128), (20, 120, 160, 6, 256), (16, 64, 96, 5, 32), (33, 48, 64, 2, 64), (24, 90, 120, 7, 100)])]
    one, st1, _ = wl.DeviceRunner(dev, budget_bytes=1e12).run_list(specs)
    many, st2, _ = wl.DeviceRunner(dev, budget_bytes=1.0).run_list(specs)
    assert st1["chunks"] == 1 and st2["chunks"] == len(specs)
    assert [r["digest"] for r in one] == [r["digest"] for r in many]
    assert sum(r["status"] == 1 for r in one) >= 4 and st1["k2_tiles"] == st2["k2_tiles"] > 0
    # and the digests are those of the plain in-memory entry point on the same scenes
    from s2d_b200.pipeline import Batch, VideoInput
    from s2d_b200.synth import make_scene_device
    for s, r in zip(specs, one):
        sc = make_scene_device(s.seed, s.T, s.H, s.W, s.M, s.P, dev)
        b = Batch([VideoInput(sc["labels"], sc["tracks"], sc["vis"], max_label=s.M)])
        b.run()
        torch.cuda.synchronize()
        assert wl.video_digest(b.fetch_summary(), 0, 0, b.host_descs[0].Nm) == r["digest"]


@pytest.mark.gpu
@pytest.mark.parametrize("window", [48, 64])
def test_c3_shaped_windowed_video_vs_oracle(window):
    """BASELINE.json configs[2] down-scaled in pixels and points only: 300 frames (10 words per frame-bit row), bit-packed
    visibility flags, tracks stored for a window of <= 64 frames per query chosen on the device from stage B's output,
    two-phase run (stages A-B, windowed tracker stand-in, stage D). The oracle sees full-length tracks that are NaN
    outside every query's stored window: every field must agree."""
    from oracle import keymask_oracle as ko
    from tests.test_gpu_parity import _assert_same_as_oracle
    dev = torch.device("cuda:0")
    spec = wl.VideoSpec("c3small", 31000 + window, 300, 90, 160, 6, 64, window=window, vis_bits=True)
    res, st, (batch, scenes, vids) = wl.DeviceRunner(dev).run_list([spec], keep_batch=True)
    got = batch.decode()[0]
    sc = scenes[0]
    Nm, Ttr, P, _ = sc["tracks"].shape
    assert Ttr == window and Nm > 1200
    ts = sc["tstart"].cpu().numpy()
    win = sc["tracks"].cpu().numpy()
    full = np.full((Nm, spec.T, P, 2), np.nan, np.float32)
    for q in range(Nm):
        full[q, ts[q]:ts[q] + window] = win[q]
    words = sc["vis"].cpu().numpy().view(np.uint32)
    vis = np.unpackbits(words.view(np.uint8).reshape(Nm, spec.T, -1), axis=2, bitorder="little")[:, :, :P]
    ref = ko.discover(sc["labels"].cpu().numpy(), full, vis)
    _assert_same_as_oracle(got, ref, 1000)
    assert res[0]["status"] == 1 and res[0]["candidates"] == len(ref["queries"])
    # windows are really partial: most queries' [v0, v1] is the whole video
    assert np.mean([q["v_range"][1] - q["v_range"][0] + 1 for q in ref["queries"]]) > 2 * window
