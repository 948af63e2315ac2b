"""Where does the end-to-end step differ from the device-resident step?  Times, with CUDA events on
the main stream: run_device on resident input, the pipelined run_stream loop, and run_stream with
the H2D copy replaced by a device-to-device copy."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402
from rtpe_b200 import inference  # noqa: E402

PARSER_KW = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
                 use_detection_val=True, ignore_too_much=False, tag_per_joint=True, nms_ksize=5,
                 nms_padding=2)


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def main():
    chunk = int(os.environ.get("CHUNK", "64"))
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None).cuda()
    model[1].chunk_size = chunk
    model[1].freeze()
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)
    xh = torch.randn(32, 3, 640, 640).pin_memory()
    xd = xh.cuda()

    def dev(n):
        for _ in range(n):
            pipe.run_device(xd, True, True)

    def stream_host(n):
        for out in pipe.run_stream((xh for _ in range(n)), True, True):
            pass

    def stream_dev(n):
        for out in pipe.run_stream((xd for _ in range(n)), True, True):
            pass

    def serial_host(n):
        for _ in range(n):
            pipe.run_device(xh.cuda(non_blocking=True), True, True)

    for name, fn in (("run_device (resident)", dev), ("run_stream (pinned host)", stream_host),
                     ("run_stream (device source)", stream_dev), ("serial H2D + run_device", serial_host)):
        fn(3)
        for rep in range(2):
            ms, wall = timed(fn, 8)
            print("%-30s %8.3f ms/step (events)  %8.3f ms/step (wall)" % (name, ms, wall))


if __name__ == "__main__":
    main()
