"""Phase breakdown of one bench step (flip-test forward + aggregation + decode), CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402
from rtpe_b200 import inference  # noqa: E402

PARSER_KW = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
                 use_detection_val=True, ignore_too_much=False, tag_per_joint=True, nms_ksize=5,
                 nms_padding=2)


class Timer:
    def __init__(self):
        self.marks = []

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks.append((name, e))

    def report(self, title):
        torch.cuda.synchronize()
        tot = self.marks[0][1].elapsed_time(self.marks[-1][1])
        print("%s: total %.3f ms" % (title, tot))
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            print("   %-28s %8.3f ms" % (n1, e0.elapsed_time(e1)))


def main():
    batch = int(os.environ.get("BATCH", "32"))
    chunk = int(os.environ.get("CHUNK", "16"))
    size = int(os.environ.get("SIZE", "640"))
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None).cuda()
    net = model[1]
    net.chunk_size = chunk
    net.freeze()
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)
    x = torch.randn(batch, 3, size, size, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            pipe.run_device(x)
        torch.cuda.synchronize()
        for rep in range(2):
            t = Timer()
            t.mark("start")
            both = torch.cat((x, torch.flip(x, [3])), 0)
            t.mark("cat+flip")
            y0, y1 = model(both)
            t.mark("forward (%d fwd)" % (2 * batch))
            n = batch
            det, tag = inference.aggregate_scale(y0[:n], y1[:n], y0[n:], y1[n:], (size, size), 17)
            t.mark("aggregate")
            val_k, ind_k, _, tag_k = parser.top_k_device(det, tag)
            t.mark("top_k")
            ans, count, pmax = parser.match_device(val_k, ind_k, tag_k, size)
            t.mark("match (+overflow sync)")
            parser.adjust_device(ans, count, det)
            t.mark("adjust")
            parser.refine_device(det, tag, ans, count)
            t.mark("refine")
            a2, c2, s2 = inference.pad_results(ans, count, torch.zeros_like(ans[:, :, 0, 0]),
                                               parser.person_capacity)
            t.mark("pad")
            t.report("step rep %d (batch %d, chunk %d)" % (rep, batch, chunk))
        # whole step as the bench runs it
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            pipe.run_device(x)
        e1.record()
        torch.cuda.synchronize()
        print("run_device: %.3f ms/step -> %.1f img/s" % (e0.elapsed_time(e1) / 5, batch * 5 / e0.elapsed_time(e1) * 1e3))
        # forward pieces
        t = Timer()
        plan = net._get_plan(chunk, size, size, net._mode(), x.device, torch.float16)
        xh = both[:chunk].half()
        t.mark("start")
        xh2 = both.half()
        t.mark("tofp16 (all %d)" % (2 * batch))
        plan.in_buf.copy_(xh)
        t.mark("copy in (1 chunk)")
        import rtpe_b200._lib as L
        L.check(L.load().brtpe_plan_graph_launch(plan.handle, L.stream_ptr(x.device)), "graph")
        t.mark("graph (1 chunk)")
        r = [o.clone() for o in plan.outs]
        t.mark("copy out (1 chunk)")
        r2 = [o.float() for o in (y0, y1)]
        t.mark("tofp32 (all)")
        t.report("forward pieces")
        print("people/img", float(count.float().mean()))


if __name__ == "__main__":
    main()
