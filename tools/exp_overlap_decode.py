"""Experiment: decode of batch i on a side stream while the network graph of batch i+1 runs.
Prints ms per step for the serial schedule and the overlapped one (same kernels, same work)."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402
from rtpe_b200 import inference, _lib as L  # noqa: E402

PARSER_KW = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
                 use_detection_val=True, ignore_too_much=False, tag_per_joint=True, nms_ksize=5,
                 nms_padding=2)


def main():
    batch, size, steps = int(os.environ.get("BATCH", "32")), 640, int(os.environ.get("STEPS", "10"))
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None).to(dev)
    net = model[1]
    net.chunk_size = 64
    net.freeze()
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 3, size, size, generator=g).to(dev)
    lib = L.load()

    def decode_nosync(det, tag):
        n, j, h, w = det.shape
        val_k, ind_k, _, tag_k = parser.top_k_device(det, tag)
        ans, count, pmax, flag = parser.match_device(val_k, ind_k, tag_k, w, defer_overflow=True)
        parser.adjust_device(ans, count, det)
        scores = torch.empty((n, pmax), dtype=torch.float32, device=dev)
        L.check(lib.brtpe_scores(L.ptr(ans), L.ptr(count), L.ptr(scores), n, j, ans.shape[3] - 3, pmax,
                                 L.stream_ptr(dev)), "brtpe_scores")
        parser.refine_device(det, tag, ans, count)
        return ans, count, scores, flag

    def serial(k):
        for _ in range(k):
            det, tag = pipe.forward_aggregate(x)
            out = decode_nosync(det, tag)
        return out

    side = torch.cuda.Stream(device=dev)
    keep = collections.deque(maxlen=3)

    def overlapped(k):
        main = torch.cuda.current_stream(dev)
        out = None
        for _ in range(k):
            det, tag = pipe.forward_aggregate(x)
            ev = torch.cuda.Event()
            ev.record(main)
            det.record_stream(side)
            tag.record_stream(side)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                out = decode_nosync(det, tag)
            keep.append((det, tag, out))
        main.wait_stream(side)
        return out

    def timed(fn, label):
        fn(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print("%-12s %.3f ms / step  %.1f images/s  (people/img %.1f)" % (
            label, ms, batch / ms * 1e3, float(out[1].float().mean())), flush=True)
        return out

    a = timed(serial, "serial")
    b = timed(overlapped, "overlapped")
    a2 = timed(serial, "serial")
    b2 = timed(overlapped, "overlapped")
    print("identical results:", all(torch.equal(p, q) for p, q in zip(a[:3], b[:3])))


if __name__ == "__main__":
    main()
