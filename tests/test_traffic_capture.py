"""bench.py reports `roofline.traffic` from profiles/k2_traffic.json (DRAM bytes of the C2 votes launch from a committed
`ncu --set full` capture) while s2d_version() equals the capture's lib_version. This test backs that rule: as long as the
version is unchanged, the benchmarked kernel in the built library must compile to the same SASS as at the commit the capture
was taken on (tools/sass_same.sh). Needs nvcc, cuobjdump, the built library and the git history; skipped otherwise."""
import json
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_capture_matches_the_built_kernel():
    tj = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
    src = open(os.path.join(ROOT, "s2d_b200", "csrc", "capi.cu")).read()
    version = int(re.search(r"s2d_version\(void\)\s*\{\s*return\s+(\d+)\s*;", src).group(1))
    if version != int(tj["lib_version"]):
        pytest.skip("library version differs from the capture's: bench.py reports no traffic figure")
    lib = os.path.join(ROOT, "s2d_b200", "libs2d_b200.so")
    if not (shutil.which("nvcc") and shutil.which("cuobjdump") and os.path.exists(lib)):
        pytest.skip("needs nvcc, cuobjdump and the built library")
    commit = tj["captured_at_commit"]
    if subprocess.run(["git", "-C", ROOT, "cat-file", "-e", commit + "^{commit}"], capture_output=True).returncode != 0:
        pytest.skip("git history without the capture's commit")
    r = subprocess.run([os.path.join(ROOT, "tools", "sass_same.sh"), commit, tj["kernel_symbol"]], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("SAME"), (r.stdout, r.stderr[-500:])
