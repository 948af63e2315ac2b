"""Generate tests/golden/config1_bundled.npz: BASELINE.json configs[0] run with the UNMODIFIED
reference (/root/reference) in this container.

    python -m oracle.make_golden_config1

Per bundled image (data/000000001000.jpg, data/000000002685.jpg) the reference's own loop body
(validate_hhrnet.py:84-105) runs on the CPU in float32:

    PIL decode -> resize_align_multi_scale(np.array(img), 640, 1, 1) (transforms.py:155-192)
    -> ToTensor + Normalize (validate_hhrnet.py:63-67) -> PoseHigherResolutionNet (default
    PyTorch init under torch.manual_seed(0); helpers.py:37-67 hyper-parameters are the class
    defaults) -> bilinear align_corners=True to the original (h, w) (validate_hhrnet.py:94-98)
    -> HeatmapParser.parse(adjust=True, refine=True) with validate_hhrnet.py:40-47 parameters.

The reference's ``init_weights`` is NOT called: its N(0, 0.001) weights give ~0 heat-maps and an
empty decode.  ``munkres`` is satisfied by oracle/munkres_ref.py (see its header).

Stored (the full maps are 20-50 MB each; the GPU box re-computes them with the oracle port, which
this fixture pins): the JPEG bytes, the network input's shape and a strided sample of it, strided
samples + max|.| of y0 / y1 / hms / aes, and the complete parse() result.
"""
from __future__ import annotations

import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import REF_ROOT, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "config1_bundled.npz")
IMAGES = ["000000001000.jpg", "000000002685.jpg"]
STRIDE = 4           # spatial sampling stride of the stored network outputs
STRIDE_BIG = 8       # ... of the stored input / full-resolution maps
HM_PARSER_PARAMS = {"max_num_people": 30, "detection_threshold": 0.1, "tag_threshold": 1.0,
                    "use_detection_val": True, "ignore_too_much": False, "tag_per_joint": True,
                    "nms_ksize": 5, "nms_padding": 2}


def main():
    from PIL import Image
    import torchvision
    ref_group, ref_model = load_reference()
    import rtpe.third_party.transforms as ref_tf
    tf = torchvision.transforms.Compose([
        torchvision.transforms.ToTensor(),
        torchvision.transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225],
                                         inplace=True)])
    torch.manual_seed(0)
    net = ref_model.PoseHigherResolutionNet().eval()
    parser = ref_group.HeatmapParser(num_joints=17, **HM_PARSER_PARAMS)
    out = {"stride": np.int64(STRIDE), "stride_big": np.int64(STRIDE_BIG)}
    for k, name in enumerate(IMAGES):
        with open(os.path.join(REF_ROOT, "data", name), "rb") as f:
            raw = f.read()
        img = Image.open(io.BytesIO(raw)).convert("RGB")
        resized, center, scale = ref_tf.resize_align_multi_scale(np.array(img), 640, 1, 1)
        t = tf(resized).unsqueeze(0)
        w, h = img.size
        with torch.no_grad():
            preds, refined = net(t)
            hms = torch.nn.functional.interpolate(refined, (h, w), mode="bilinear", align_corners=True)
            aes = torch.nn.functional.interpolate(preds[:, 17:], (h, w), mode="bilinear",
                                                  align_corners=True)
        grouped, scores = parser.parse(hms, aes.unsqueeze(-1), adjust=True, refine=True)
        people = np.asarray(grouped[0], np.float32)
        p = "img%d_" % k
        out[p + "jpeg"] = np.frombuffer(raw, np.uint8)
        out[p + "hw"] = np.asarray([h, w], np.int64)
        out[p + "input_shape"] = np.asarray(t.shape, np.int64)
        out[p + "input_s"] = t[:, :, ::STRIDE_BIG, ::STRIDE_BIG].numpy()
        for nme, v in (("y0", preds), ("y1", refined), ("hms", hms), ("aes", aes)):
            st = STRIDE if nme in ("y0", "y1") else STRIDE_BIG
            out[p + nme + "_s"] = v[:, :, ::st, ::st].numpy()
            out[p + nme + "_absmax"] = np.float64(v.abs().max().item())
        out[p + "people"] = people
        out[p + "scores"] = np.asarray(scores, np.float32)
        print(name, "input", tuple(t.shape), "orig", (h, w), "people", people.shape,
              "hm range", float(hms.min()), float(hms.max()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
