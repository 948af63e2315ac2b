"""Aggregation + decode kernels alone on bench-shaped inputs (batch B, 640x640, flip): timing
per kernel with CUDA events; used under ncu for the HBM-bound kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402
from rtpe_b200 import inference  # noqa: E402


def main():
    b = int(os.environ.get("BATCH", "32"))
    s = int(os.environ.get("SIZE", "640"))
    reps = int(os.environ.get("REPS", "5"))
    g = torch.Generator().manual_seed(0)
    y0 = (torch.randn((2 * b, 34, s // 4, s // 4), generator=g) * 0.2).cuda()
    y1 = (torch.randn((2 * b, 17, s // 2, s // 2), generator=g) * 0.2).cuda()
    parser = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for r in range(reps):
        ev[0].record()
        det, tag = inference.aggregate_scale(y0[:b], y1[:b], y0[b:], y1[b:], (s, s), 17)
        ev[1].record()
        val_k, ind_k, _, tag_k = parser.top_k_device(det, tag)
        ev[2].record()
        ans, count, _ = parser.match_device(val_k, ind_k, tag_k, s)
        ev[3].record()
        parser.adjust_device(ans, count, det)
        ev[4].record()
        parser.refine_device(det, tag, ans, count)
        ev[5].record()
        torch.cuda.synchronize()
    gb = 1e-9
    bytes_agg = det.numel() * 4 + tag.numel() * 4 + (y0.numel() + y1.numel()) * 4
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
    print("aggregate %.3f ms  %.0f GB/s (reads+writes %.2f GB)" % (ms[0], bytes_agg * gb / ms[0] * 1e3, bytes_agg * gb))
    print("top_k     %.3f ms  %.0f GB/s" % (ms[1], det.numel() * 4 * gb / ms[1] * 1e3))
    print("match     %.3f ms" % ms[2])
    print("adjust    %.3f ms" % ms[3])
    print("refine    %.3f ms  %.0f GB/s" % (ms[4], (det.numel() + tag.numel()) * 4 * gb / ms[4] * 1e3))
    print("people/img", float(count.float().mean()))


if __name__ == "__main__":
    main()
