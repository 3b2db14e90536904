"""TEST INFRASTRUCTURE ONLY (build container: needs /root/reference). Runs the unmodified reference's
convert_lblimg_to_maskid (keymask_ident/crw_utils.py:688-711, the rule load_masks applies to every colour-coded mask
frame) on seeded colour images and stores inputs + outputs in tests/golden_cpu/colors.npz, the fixture
tests/test_oracle_golden.py::test_colour_to_label_rule_vs_reference_golden checks the product's host rule against.

    python -m oracle.make_golden_colors
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402


def images():
    rng = np.random.default_rng(77)
    out = []
    for (H, W, ncol, black) in [(24, 31, 5, True), (40, 56, 40, True), (17, 23, 7, False), (9, 11, 70, True), (8, 8, 1, False)]:
        pal = rng.integers(0, 256, size=(ncol, 3)).astype(np.uint8)
        pal[pal.sum(1) == 0] = 1
        if ncol >= 5:      # tuples whose lexicographic order differs from the order of their sums / last channels
            pal[:4] = [[1, 255, 0], [2, 0, 0], [1, 0, 255], [0, 255, 255]]
        idx = rng.integers(0, ncol + (1 if black else 0), size=(H, W))
        img = np.zeros((H, W, 3), np.uint8)
        img[idx < ncol] = pal[idx[idx < ncol]]
        out.append(img)
    return out


def main():
    ref_harness._install_stubs(lambda checkpoint=None: None)
    if ref_harness.REFERENCE_DIR not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_DIR)
    sys.modules.pop("crw_utils", None)
    import crw_utils                                      # the unmodified reference module
    data = {}
    for i, img in enumerate(images()):
        ids = np.asarray(crw_utils.convert_lblimg_to_maskid(img))
        data[f"rgb{i}"] = img
        data[f"ids{i}"] = ids.astype(np.int64)
        print(i, img.shape, "labels", int(ids.max()))
    d = os.path.join(ROOT, "tests", "golden_cpu")
    os.makedirs(d, exist_ok=True)
    np.savez_compressed(os.path.join(d, "colors.npz"), **data)


if __name__ == "__main__":
    main()
