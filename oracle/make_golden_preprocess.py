"""Generate tests/golden/preprocess_*.npz by running the UNMODIFIED reference pre-processing
(/root/reference rtpe/third_party/transforms.py + cv2 + torchvision, teacher_inference.py:70-79)
in this container, and check oracle/preprocess_ref.py against it bit for bit -- on small synthetic
images (committed as fixtures) and on the two bundled JPEGs at full size (checked here only; the
JPEGs do not travel to the GPU box).

    python -m oracle.make_golden_preprocess
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import cv2  # noqa: E402,F401
import torch  # noqa: E402
import torchvision  # noqa: E402
from rtpe.third_party import transforms as RT  # noqa: E402

from oracle import preprocess_ref as P  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_chain(img, input_size, cur, mn):
    resized, center, scale = RT.resize_align_multi_scale(img, input_size, cur, mn)
    pre = torchvision.transforms.Compose([
        torchvision.transforms.ToTensor(),
        torchvision.transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return resized, center, scale, pre(resized).numpy()


def check(img, input_size, cur, mn, what):
    resized, center, scale, t = reference_chain(img, input_size, cur, mn)
    o_resized, o_center, o_scale = P.resize_align_multi_scale(img, input_size, cur, mn)
    assert resized.shape == o_resized.shape, (what, resized.shape, o_resized.shape)
    assert np.array_equal(center, o_center) and np.array_equal(scale, o_scale), what
    nd = int((resized != o_resized).sum())
    assert nd == 0, "%s: %d of %d bytes differ" % (what, nd, resized.size)
    assert np.array_equal(t, P.to_tensor_normalize(o_resized)), what
    return resized, center, scale, t


def main():
    rng = np.random.default_rng(7)
    cases = {}
    for k, (h, w, size, cur, mn) in enumerate([(48, 64, 128, 1, 1), (75, 50, 128, 1, 1),
                                               (60, 60, 128, 1, 1), (41, 97, 64, 2.0, 0.5),
                                               (64, 48, 128, 0.5, 0.5), (33, 71, 64, 1, 1)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        # smooth half of the cases: real images are not white noise
        if k % 2:
            img = cv2.GaussianBlur(img, (7, 7), 2.0)
        resized, center, scale, t = check(img, size, cur, mn, "synthetic %d" % k)
        cases["img_%d" % k] = img
        cases["args_%d" % k] = np.array([size, cur, mn], np.float64)
        cases["resized_%d" % k] = resized
        cases["center_%d" % k] = center
        cases["scale_%d" % k] = scale
        if k in (0, 5):                       # float32 tensors are 4x the bytes: two cases only
            cases["tensor_%d" % k] = t
    # post-processing (transforms.py:50-56, :195-202): person arrays back to image coordinates
    for k, (h, w, size) in enumerate([(480, 640, 640), (555, 640, 640), (640, 427, 512)]):
        dsize, center, scale = RT.get_multi_scale_size(np.zeros((h, w, 3), np.uint8), size, 1, 1)
        persons = [rng.uniform(0, min(dsize), (17, 5)).astype(np.float32) for _ in range(3)]
        final = RT.get_final_preds([persons], center, scale, [dsize[0], dsize[1]])
        cases["post_hw_%d" % k] = np.array([h, w, size])
        cases["post_in_%d" % k] = np.stack(persons)
        cases["post_out_%d" % k] = np.stack(final)
    cases["n_post"] = np.array(3)
    cases["n"] = np.array(6)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "preprocess_cases.npz"), **cases)
    print("wrote preprocess_cases.npz", sum(v.nbytes for v in cases.values()) // 1024, "KiB raw")
    # wider sweep of sizes (not stored) + the bundled JPEGs at full size
    for _ in range(40):
        h, w = int(rng.integers(20, 300)), int(rng.integers(20, 300))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        check(img, int(rng.choice([64, 128, 256])), 1, 1, "sweep %dx%d" % (h, w))
    from PIL import Image
    for name in ("000000001000.jpg", "000000002685.jpg"):
        img = np.array(Image.open(os.path.join("/root/reference/data", name)).convert("RGB"))
        for cur, mn in ((1, 1), (2.0, 0.5), (0.5, 0.5)):
            r, *_ = check(img, 640, cur, mn, name)
            print(name, img.shape, "->", r.shape, "scale", cur, "bit-exact")


if __name__ == "__main__":
    main()
