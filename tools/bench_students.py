"""Forward throughput of the student drop-ins on one B200: batch 64 synthetic
512 x 512 (BASELINE config 4's input), bf16 and fp32, CUDA-event timing over plan replays, plus the
per-launch profile of the bf16 plan (which layers the time goes to)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402
from rtpe_b200 import _lib as L  # noqa: E402


def build(name, half):
    torch.manual_seed(0)
    if name == "CamStudent":
        return rtpe_b200.CamStudent(None, "cuda", half_precision=half), {}
    if name == "RefinerStudent":
        return rtpe_b200.RefinerStudent(None, "cuda", half_precision=half), {}
    if name == "MultistageStudent":
        return rtpe_b200.MultistageStudent(None, "cuda", half_precision=half), {}
    if name == "AttentionStudentSteps":
        return rtpe_b200.AttentionStudentSteps(None, "cuda", 80, 17, 0, half), {"alt": True}
    if name == "AttentionStudent":
        return rtpe_b200.AttentionStudent(None, "cuda", inplanes=48, half_precision=half), {}
    raise ValueError(name)


def run(name, half, batch, size, reps):
    net, opt = build(name, half)
    net = net.eval()
    net.chunk_size = batch
    net.freeze()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 3, size, size, generator=g).cuda()
    kw = {"alt": torch.randn(batch, 3, size, size, generator=g).cuda()} if opt.get("alt") else {}
    with torch.no_grad():
        for _ in range(3):
            net(x, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            net(x, **kw)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    plan = next(iter(net._plans.values()))
    out = {"mode": "bf16" if half else "fp32", "forward_ms": ms, "images_per_s": batch / ms * 1e3,
           "conv_gflop_per_image": plan.conv_flops / batch / 1e9,
           "forward_tflops": plan.conv_flops / ms / 1e9, "launches": plan.num_ops}
    if half:
        tms, kinds, fl = net.plan_profile(batch, size, size)
        names = {0: "conv_umma", 1: "conv_ffma", 2: "other", 3: "conv_halo"}
        tot = sum(tms)
        share = {}
        for t, k in zip(tms, kinds):
            share[names[k]] = share.get(names[k], 0.0) + t
        out["profile_ms"] = {k: round(v, 3) for k, v in share.items()}
        aux_names = {1: "avgpool", 2: "se_partial", 3: "se_gate", 4: "cam_mix", 5: "att_add", 6: "resize_nhwc",
                     7: "image_nhwc", 8: "space_to_depth", 9: "att_mul"}
        other = {}
        for t, k, op in zip(tms, kinds, plan.recorder.ops):
            if k == 2:
                label = aux_names.get(op[1], "aux") if op[0] == "aux" else op[0]
                other[label] = other.get(label, 0.0) + t
        out["other_ms"] = {k: round(v, 3) for k, v in sorted(other.items(), key=lambda kv: -kv[1])}
        out["profile_total_ms"] = round(tot, 3)
    return out


def main():
    batch, size = int(os.environ.get("BATCH", "64")), int(os.environ.get("SIZE", "512"))
    names = sys.argv[1:] or ["CamStudent", "RefinerStudent", "MultistageStudent", "AttentionStudentSteps"]
    out = {"workload": "student forward, batch %d, %dx%d synthetic, random-init weights" % (batch, size, size)}
    for name in names:
        out[name] = {"bf16": run(name, True, batch, size, 5), "fp32": run(name, False, batch, size, 2)}
        print(name, json.dumps(out[name]), file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
