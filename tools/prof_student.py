import sys, os
sys.path.insert(0, "/root/repo")
import torch, rtpe_b200
torch.manual_seed(0)
net = rtpe_b200.AttentionStudent(None, "cuda", inplanes=48, num_heatmaps=17, ae_dims=1, half_precision=True).eval()
net.chunk_size = 64
x = torch.randn(64, 3, 512, 512).cuda()
with torch.no_grad():
    net(x)
net.plan_profile(64, 512, 512, torch.float32)
ms, kinds, fl = net.plan_profile(64, 512, 512, torch.float32)
plan = net._get_plan(64, 512, 512, "bf16", x.device, torch.float32)
ops = plan.recorder.ops if hasattr(plan, "recorder") else None
print("total", sum(ms), "ops", len(ms))
rows = []
for i, (m, k, f) in enumerate(zip(ms, kinds, fl)):
    desc = ""
    if ops is not None and i < len(ops):
        op = ops[i]
        if op[0] == "conv":
            d = op[1]
            desc = "conv %dx%d %d->%d taps%d s%d eng%d" % (d.Hin, d.Win, d.Cin, d.Cout, d.ntaps, d.in_stride, d.engine)
        else:
            desc = op[0] + (" kind%d" % op[1] if op[0] == "aux" else "")
    rows.append((m, k, f, desc))
for m, k, f, desc in sorted(rows, reverse=True)[:25]:
    print("%8.3f ms kind %d %8.1f TF  %s" % (m, k, f / max(m, 1e-9) / 1e9, desc))
