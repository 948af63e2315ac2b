"""Config 3 of BASELINE.json on one GPU's shard: multi-scale (2.0 / 1.0 / 0.5) + flip-test aggregation
to 640 x 640 + AE grouping for a shard of the 256-image batch (32 images = the 8-GPU shard).  Synthetic
inputs at 1280^2 / 640^2 / 320^2 (transforms.py:155-176 with min_scale 0.5), random-init W48, bf16."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402
from rtpe_b200 import inference  # noqa: E402


def main():
    shard = int(os.environ.get("SHARD", "32"))
    sub = int(os.environ.get("SUB", "32"))         # images per multi-scale pass (1280^2 activations are 4x)
    reps = int(os.environ.get("REPS", "3"))
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None).cuda()
    model[1].chunk_size = 2 * sub
    model[1].freeze()
    parser = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False, nms_ksize=5, nms_padding=2)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)
    g = torch.Generator().manual_seed(1)
    xs = {s: torch.randn(sub, 3, int(640 * s), int(640 * s), generator=g).cuda() for s in (2.0, 1.0, 0.5)}

    def one_pass():
        det, tag = pipe.forward_aggregate_multiscale([(s, xs[s]) for s in (2.0, 1.0, 0.5)], (640, 640))
        return parser.decode_device(det, tag)

    for _ in range(2):
        one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for _ in range(shard // sub):
            ans, count, scores = one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2 * (4 + 1 + 0.25) * 298.94e9 * shard
    print(json.dumps({"workload": "config 3: multi-scale 2.0/1.0/0.5 + flip, shard of %d images, %d per pass" % (shard, sub),
                      "ms_per_shard": ms, "images_per_s": shard / ms * 1e3,
                      "forward_tflops_effective": flop / ms / 1e9,
                      "people_per_image": float(count.float().mean()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
