"""bench.py's reference arm runs on the host alone (no GPU, no /root/reference) and prints one JSON line with the
keys the driver reads; the N>1 launch lets rank 0 print and the other ranks exit 0 without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    if extra_env:
        env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "keymask_discovery_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("tiny")


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_ignores_torchrun_thread_cap():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use every host core, or the per-N
    ratios of the scaling run are not comparable (round-1 verdict)."""
    r = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    # both arms print the same `config` (bench.make_config)
    assert set(d["config"]) == {"workload", "videos_per_gpu", "frames_per_video", "resolution", "masks_per_frame",
                                "tracks_per_query", "queries_per_video", "point_order", "partition", "cache"}


def test_reference_arm_of_a_list_workload():
    """--impl reference --workload c4: the CPU port on one query of a few videos of the list (down-scaled here), same
    contract line, strong scaling."""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c4", "--list-videos", "4",
                        "--list-scale", "0.1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["scaling"] == "strong" and d["value"] > 0
    assert d["config"]["workload"].startswith("c4") and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
