"""Quick GPU timing of the forward plan (per-op device times) and of the decode kernels."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402


def main():
    chunk = int(os.environ.get("CHUNK", "8"))
    size = int(os.environ.get("SIZE", "640"))
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None).cuda()
    net = model[1]
    net.chunk_size = chunk
    x = torch.randn(chunk, 3, size, size, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            y = model(x)
        torch.cuda.synchronize()
        t0 = time.time()
        iters = 10
        for _ in range(iters):
            y = model(x)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / iters
    flops = 298.94e9 * (size / 640.0) ** 2 * chunk
    print("forward bf16 chunk=%d %dx%d: %.3f ms/chunk, %.1f img/s, %.1f TFLOP/s" %
          (chunk, size, size, dt * 1e3, chunk / dt, flops / dt / 1e12))
    ms, kinds, fl = net.plan_profile(chunk, size, size, torch.float16)
    tot = sum(ms)
    umma = sum(m for m, k in zip(ms, kinds) if k in (0, 3))
    ffma = sum(m for m, k in zip(ms, kinds) if k == 1)
    other = sum(m for m, k in zip(ms, kinds) if k == 2)
    uf = sum(f for f, k in zip(fl, kinds) if k in (0, 3))
    print("profile: total %.3f ms over %d ops; umma %.3f ms (%.1f TFLOP/s), ffma %.3f ms, other %.3f ms"
          % (tot, len(ms), umma, uf / max(umma, 1e-9) / 1e9, ffma, other))
    plan = net._get_plan(chunk, size, size, net._mode(), x.device, torch.float16)
    groups = {}
    for i, op in enumerate(plan.recorder.ops):
        if op[0] == "conv":
            d = op[1]
            key = ("conv k%d %dx%d s%d %d->%d @%dx%d" % (kinds[i], 3 if d.ntaps == 9 else (2 if d.ntaps == 4 else 1),
                   3 if d.ntaps == 9 else (2 if d.ntaps == 4 else 1), d.in_stride, d.Cin, d.Cout, d.Hm, d.Wm))
        else:
            key = op[0]
        g = groups.setdefault(key, [0, 0.0, 0.0])
        g[0] += 1; g[1] += ms[i]; g[2] += fl[i]
    print("%-44s %5s %9s %7s %9s" % ("op group", "count", "ms", "%time", "TFLOP/s"))
    for key, (cnt, m, f) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %5d %9.4f %6.1f%% %9.1f" % (key, cnt, m, 100 * m / tot, f / max(m, 1e-9) / 1e9))

    # decode stress (config 5 shape, smaller batch)
    nb = int(os.environ.get("DECODE_BATCH", "128"))
    det, tag = rtpe_b200.synth_decode_batch(nb, height=320, width=320, device="cuda", seed=1234)
    hp = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False)
    for _ in range(2):
        hp.decode_device(det, tag)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    val_k, ind_k, _, tag_k = hp.top_k_device(det, tag)
    ev[1].record()
    ans, count, pmax = hp.match_device(val_k, ind_k, tag_k, 320)
    ev[2].record()
    hp.adjust_device(ans, count, det)
    ev[3].record()
    hp.refine_device(det, tag, ans, count)
    ev[4].record()
    torch.cuda.synchronize()
    names = ["top_k", "match", "adjust", "refine"]
    for i, nme in enumerate(names):
        print("decode %-7s %.3f ms for %d images" % (nme, ev[i].elapsed_time(ev[i + 1]), nb))
    bytes_topk = nb * 17 * 320 * 320 * 4
    print("top_k achieved %.1f GB/s (det read once)" % (bytes_topk / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9))
    print("refine achieved %.1f GB/s (det+tag algorithmic)" % (2 * bytes_topk / (ev[3].elapsed_time(ev[4]) * 1e-3) / 1e9))
    print("people/img mean", float(count.float().mean()))


if __name__ == "__main__":
    main()
