"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz + *.json by running the UNMODIFIED
reference (/root/reference/keymask_ident, via oracle/ref_harness.py) on small seeded scenes.

    python -m oracle.make_golden            # regenerate every case (build container only)

Each case stores its inputs (labels, tracks, vis) and the reference's outputs so that the
GPU box - where /root/reference does not exist - can run the parity tests from the fixtures
alone. Cases cover the edge behaviour listed in SURVEY.md section 4 / Appendix A.7.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from s2d_b200.synth import make_scene  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _mut_nocand(sc):
    # six rows never visible -> their own DBSCAN cluster with no majority run -> no candidates
    sc.vis[:6] = 0


def _mut_many_clusters(sc):
    # 11 visibility patterns of 8 rows each: pseudo-random half-dense codewords (pairwise hamming
    # far above 0.2*T) that are all-ones on the group's own 4 frames, so every cluster has
    # candidates and cluster_10 sorts before cluster_2 (Appendix A.7 quirk 3)
    T = sc.vis.shape[1]
    rng = np.random.default_rng(99)
    pats = (rng.random((11, T)) < 0.5).astype(np.uint8)
    for g in range(11):
        pats[g, 4 * g:4 * g + 4] = 1
    d = (pats[:, None, :] != pats[None, :, :]).sum(-1)
    assert (d[~np.eye(11, dtype=bool)] > 0.2 * T).all()
    for q in range(sc.vis.shape[0]):
        g = min(int(sc.query_frame[q]) // 4, 10)
        sc.vis[q] = pats[g][:, None]


def _mut_one2x(sc):
    # queries of label 2 get half of label 3's points of the same frame -> two masks > 0.25 each
    lut = {(int(f), int(l)): q for q, (f, l) in enumerate(zip(sc.query_frame, sc.query_label))}
    P = sc.tracks.shape[2]
    for (f, l), q in lut.items():
        if l == 2 and (f, 3) in lut:
            sc.tracks[q, :, P // 2:] = sc.tracks[lut[(f, 3)], :, P // 2:]


def _mut_empty_crop(sc):
    sc.tracks[:] = np.nan


CASES = {
    # name: (make_scene kwargs, mutator, visibility_threshold, matching_threshold)
    "basic":     (dict(seed=1234, T=12, H=96, W=128, M=5, P=64, specials=True), None, 0.3, 0.5),
    "nobg":      (dict(seed=1235, T=12, H=96, W=128, M=5, P=64, full_cover_frames=(2, 5)), None, 0.3, 0.5),
    "dups":      (dict(seed=1236, T=10, H=64, W=96, M=4, P=128, dup_rate=0.3, noise=0.0), None, 0.3, 0.5),
    "tiny":      (dict(seed=1237, T=3, H=48, W=64, M=1, P=32, occlude=False), None, 0.3, 0.5),
    "nocand":    (dict(seed=1238, T=12, H=96, W=128, M=5, P=64), _mut_nocand, 0.3, 0.5),
    "many":      (dict(seed=1239, T=44, H=48, W=64, M=2, P=32, occlude=False), _mut_many_clusters, 0.3, 0.5),
    "thr":       (dict(seed=1240, T=12, H=96, W=128, M=5, P=64, noise=2.5), None, 0.5, 0.7),
    "one2x":     (dict(seed=1241, T=12, H=96, W=128, M=5, P=64, occlude=False, noise=0.3), _mut_one2x, 0.3, 0.5),
    "emptycrop": (dict(seed=1242, T=8, H=48, W=64, M=3, P=32, occlude=False), _mut_empty_crop, 0.3, 0.5),
    "medium":    (dict(seed=1243, T=16, H=120, W=160, M=8, P=128, specials=True, noise=1.0), None, 0.3, 0.5),
}


# 36- and 40-frame videos that end with status 1: more than one 32-bit word per visibility / match bit row, so
# select / group run with TW = 2 on the GPU and are compared with the reference's output (tests/golden/ like the rest).
CASES.update({
    "long": (dict(seed=1250, T=40, H=48, W=64, M=3, P=48), None, 0.3, 0.5),
    "long_thr": (dict(seed=1251, T=36, H=60, W=80, M=4, P=64, noise=1.5), None, 0.4, 0.6),
})


def build_case(name):
    kw, mut, vthr, mthr = CASES[name]
    sc = make_scene(**kw)
    if mut is not None:
        mut(sc)
    return sc, vthr, mthr


def main(names=None):
    from oracle.compare import golden_from_reference
    from oracle.ref_harness import run_reference

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in (names or list(CASES)):
        outdir = GOLDEN_DIR
        sc, vthr, mthr = build_case(name)
        with tempfile.TemporaryDirectory() as d:
            out = run_reference(sc, d, visibility_threshold=vthr, matching_threshold=mthr)
        assert out["labels_equal"], name
        g = golden_from_reference(out)
        g["visibility_threshold"] = vthr
        g["matching_threshold"] = mthr
        g["case"] = name
        np.savez_compressed(os.path.join(outdir, f"{name}.npz"), labels=sc.labels,
                            tracks=sc.tracks, vis=sc.vis)
        with open(os.path.join(outdir, f"{name}.json"), "w") as f:
            json.dump(g, f)
        nq = len(g["queries"]) if g["queries"] else 0
        print(f"{name}: status={g['status']} queries={nq} clusters={len(g['clusters'] or [])}")


if __name__ == "__main__":
    main(sys.argv[1:] or None)
