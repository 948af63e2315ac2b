"""Batch-1 latency of the reference's in-tree path (validate_hhrnet.py:84-105, config 1 shapes):
uint8 image -> GPU pre-processing -> W48 forward -> in-tree aggregation to the original image size
-> HeatmapParser.parse(adjust, refine).  Wall clock per image (host sync after parse) and CUDA-event
time per stage."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rtpe_b200  # noqa: E402
from rtpe_b200 import inference, preprocess  # noqa: E402


def main():
    h, w = int(os.environ.get("IMG_H", "480")), int(os.environ.get("IMG_W", "640"))
    reps = int(os.environ.get("REPS", "20"))
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None).cuda()
    model[1].freeze()
    parser = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False, nms_ksize=5, nms_padding=2)
    img = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
    img_pinned = torch.from_numpy(img).pin_memory()

    def one(timers=None):
        def mark(name):
            if timers is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                timers.append((name, e))
        mark("start")
        t, center, scale = preprocess.preprocess_image(img_pinned, 640)
        mark("preprocess (H2D + warp + normalise)")
        with torch.no_grad():
            y0, y1 = model(t)
        mark("forward")
        det, tag = inference.aggregate_intree(y0, y1, (h, w))
        mark("aggregate (in-tree)")
        grouped, scores = parser.parse(det, tag, True, True)
        mark("parse (incl. D2H of the people)")
        return grouped, scores

    for _ in range(5):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        grouped, scores = one()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    timers = []
    one(timers)
    torch.cuda.synchronize()
    print("image %dx%d -> network input %s, %d people" % (h, w, "640-aligned", len(grouped[0])))
    print("wall clock per image: %.3f ms (%.1f images/s at batch 1)" % (wall, 1e3 / wall))
    for (n0, e0), (n1, e1) in zip(timers[:-1], timers[1:]):
        print("   %-40s %8.3f ms" % (n1, e0.elapsed_time(e1)))


if __name__ == "__main__":
    main()
