"""Per-launch profile of the fp32 (FFMA) plan of AttentionStudent (config 4) or the W48 teacher."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import rtpe_b200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "student"
torch.manual_seed(0)
if which == "student":
    net = rtpe_b200.AttentionStudent(None, "cuda", inplanes=48, num_heatmaps=17, ae_dims=1,
                                     half_precision=False).eval()
    n, size = 64, 512
else:
    net = rtpe_b200.get_hrnet_w48_teacher(None, half=False).cuda().eval()
    n, size = 8, 640
net.chunk_size = n
x = torch.randn(n, 3, size, size).cuda()
with torch.no_grad():
    net(x)
net.plan_profile(n, size, size, torch.float32)
ms, kinds, fl = net.plan_profile(n, size, size, torch.float32)
plan = next(iter(net._plans.values()))
rows = {}
for m, k, f, op in zip(ms, kinds, fl, plan.recorder.ops):
    if op[0] == "conv":
        d = op[1]
        key = "conv %dx%d %d->%d taps%d s%d" % (d.Hin, d.Win, d.Cin, d.Cout, d.ntaps, d.in_stride)
    else:
        key = op[0]
    r = rows.setdefault(key, [0, 0.0, 0.0])
    r[0] += 1
    r[1] += m
    r[2] += f
print("total %.3f ms over %d launches" % (sum(ms), len(ms)))
for key, (c, m, f) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%-40s %3d %9.3f ms %7.1f TFLOP/s" % (key, c, m, f / max(m, 1e-9) / 1e9))
