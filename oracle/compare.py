"""TEST INFRASTRUCTURE ONLY - canonical comparison of a hot-path result (from the oracle, from
the CUDA pipeline, or from a golden fixture produced by the real reference).

A *golden* (see oracle/make_golden.py) stores what ref_harness.run_reference returned, reduced
to json-able form. `check_against_golden` asserts the integer/set quantities bit-exactly and
floats to 1e-6 relative, as BASELINE.json's north_star demands."""
from __future__ import annotations

import math

import numpy as np


def golden_from_reference(refout: dict) -> dict:
    """Reduce run_reference output to the canonical json-able golden."""
    g = {"status": refout["status"] if refout["status"] in (1, -1, None) else str(refout["status"])}
    a = refout.get("stage_a")
    if a is not None:
        rows = []
        for fr in a["video_data"]:
            for o in fr["data"]:
                rows.append({"frame_id": int(fr["frame_id"]), "object_id": int(o["object_id"]),
                             "visibility": [float(v) for v in o["visibility"]]})
        g["visibility_rows"] = rows
    g["clusters"] = refout.get("stage_b", {}).get("clusters") if refout.get("stage_b") else None
    g["candidate_files"] = refout.get("candidate_files")
    if refout.get("queries") is not None and refout.get("matches_data") is not None:
        qs = []
        for md, ql in zip(refout["matches_data"], refout["queries"]):
            qs.append({"cluster_id": md["cluster_id"], "frame_id": md["frame_id"], "mask_id": md["mask_id"],
                       "overall_mask_id": md["overall_mask_id"], "one2x": md["one2x"],
                       "matches": md["matches"], "v_range": ql["v_range"], "grid_size": ql["grid_size"],
                       "comps": [[c[0], c[1], c[2], p[0], p[1], c[3]] for c, p in zip(ql["comps"], ql["pairs"])]})
        g["queries"] = qs
    else:
        g["queries"] = None
    g["groupings"] = refout.get("groupings")
    g["one2x"] = refout.get("one2x")
    g["video_coverage_txt"] = refout.get("video_coverage_txt")
    g["cluster_coverage_txt"] = refout.get("cluster_coverage_txt")
    g["group_files"] = refout.get("group_files")
    g["tracker_calls"] = [list(c) for c in refout.get("tracker_calls", [])]
    return g


def _close(a, b, rel=1e-6):
    if isinstance(a, float) and isinstance(b, float) and (math.isnan(a) and math.isnan(b)):
        return True
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-300) or a == b


def check_against_golden(res: dict, g: dict, *, check_comps: bool = True):
    """`res` is the dict produced by oracle.keymask_oracle.discover or by
    s2d_b200.pipeline (same schema). Raises AssertionError on any difference."""
    # stage A: visibility rows, float32 exact (json round trip is exact, Appendix A.1)
    if g.get("visibility_rows") is not None:
        V = np.asarray(res["V"], np.float32)
        assert len(g["visibility_rows"]) == V.shape[0], (len(g["visibility_rows"]), V.shape)
        for i, row in enumerate(g["visibility_rows"]):
            assert int(res["query_frame"][i]) == row["frame_id"] and int(res["query_label"][i]) == row["object_id"]
            ref = np.asarray(row["visibility"], np.float32)
            assert np.array_equal(ref, V[i], equal_nan=True), (i, ref, V[i])
    # stage B
    if g.get("clusters") is not None:
        rc = res["clusters"]
        assert len(rc) == len(g["clusters"]), (len(rc), len(g["clusters"]))
        for a, b in zip(rc, g["clusters"]):
            assert a["cluster_id"] == b["cluster_id"] and a["cluster_size"] == b["cluster_size"]
            assert [list(r) for r in a["ranges"]] == [list(r) for r in b["ranges"]], (a["ranges"], b["ranges"])
            assert a["all_visible_masks"] == b["all_visible_masks"]
            assert len(a["all_candidates"]) == len(b["all_candidates"])
            for ca, cb in zip(a["all_candidates"], b["all_candidates"]):
                assert list(ca["range"]) == list(cb["range"]) and ca["candidates"] == cb["candidates"]
    # status
    gs = g["status"]
    if gs in (1, -1):
        assert res["status"] == gs, (res["status"], gs)
    if gs != 1:
        return
    # stage D: queries in processing order
    assert len(res["queries"]) == len(g["queries"])
    for a, b in zip(res["queries"], g["queries"]):
        for k in ("cluster_id", "frame_id", "mask_id", "overall_mask_id", "one2x"):
            assert int(a[k]) == int(b[k]), (k, a[k], b[k])
        assert [int(x) for x in a["matches"]] == b["matches"], (a["frame_id"], a["mask_id"])
        assert list(a["v_range"]) == list(b["v_range"])
        assert int(a["grid_size"]) == int(b["grid_size"])
        if check_comps:
            assert len(a["comps"]) == len(b["comps"])
            for ca, cb in zip(a["comps"], b["comps"]):
                assert [int(x) for x in ca[:5]] == [int(x) for x in cb[:5]], (ca, cb)
                assert _close(float(ca[5]), float(cb[5])), (ca, cb)
    # groupings
    assert len(res["groupings"]) == len(g["groupings"])
    for a, b in zip(res["groupings"], g["groupings"]):
        assert a["cluster_id"] == b["cluster_id"]
        assert a["visibility_to_temporal_factor"] == b["factor"]
        ga = {str(k): [[int(f), int(m)] for f, m in v] for k, v in a["overall_mask_ids_per_label"].items()}
        assert ga == b["groups"], (ga, b["groups"])
    # one2x + coverage
    ro, go = res["one2x"], g["one2x"]
    assert set(ro) == set(go)
    for ck in go:
        assert set(ro[ck]) == set(go[ck]), (ro[ck], go[ck])
        for gk, gv in go[ck].items():
            if gk == "avg_one2x_cluster":
                assert _close(float(ro[ck][gk]), float(gv))
            else:
                assert _close(float(ro[ck][gk]["avg_one2x"]), float(gv["avg_one2x"]))
                assert ro[ck][gk]["one2x_counts"] == gv["one2x_counts"]
                assert bool(ro[ck][gk]["noisy"]) == bool(gv["noisy"])
    assert f"Video Coverage: {res['video_coverage']:.2f}\n" == g["video_coverage_txt"]
    txts = [g["cluster_coverage_txt"][k] for k in sorted(g["cluster_coverage_txt"],
                                                          key=lambda s: int(s.split('_')[1]))]
    assert len(txts) == len(res["cluster_coverages"])
    # save_cluster_coverages (cotracker_matching.py:434-450) writes coverage i to the i-th
    # *numerically* sorted cluster dir together with factor i
    for i, txt in enumerate(txts):
        cid = sorted(int(k.split('_')[1]) for k in g["cluster_coverage_txt"])[i]
        want = (f"Cluster {cid} Coverage: {res['cluster_coverages'][i]:.2f}\n"
                f"Visibility to Temporal Factor: {res['groupings'][i]['visibility_to_temporal_factor']}\n")
        assert want == txt, (want, txt)
