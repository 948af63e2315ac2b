"""Diagnostic: AttentionStudentSteps bf16 vs oracle for each conv engine setting."""
import sys
import torch
sys.path.insert(0, ".")
from rtpe_b200 import _lib as L                                    # noqa: E402
from rtpe_b200.students import AttentionStudentSteps               # noqa: E402
from oracle.weights import fill_params_deterministic               # noqa: E402
from oracle.student_ref import attention_student_steps_forward_ref  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 80
net = AttentionStudentSteps(None, "cpu", P, 17, 1, True)
fill_params_deterministic(net, 53)
net.eval()
sd = {k: v.float() for k, v in net.state_dict().items()}
g = torch.Generator().manual_seed(54)
x = torch.randn(2, 3, 128, 128, generator=g)
alt = torch.randn(2, 3, 128, 128, generator=g)
ref = attention_student_steps_forward_ref(sd, x, alt, att_divisor=20)
net = net.cuda()
for name, eng in (("ffma", L.ENGINE_FFMA), ("umma", L.ENGINE_UMMA), ("auto", L.ENGINE_AUTO)):
    net.conv_engine = eng
    net.invalidate_plans()
    try:
        with torch.no_grad():
            got = net(x.cuda(), alt=alt.cuda(), att_divisor=20)
        print(name, [((a.double().cpu() - b.double()).abs().max() / b.double().abs().max()).item()
                     for a, b in zip(got, ref)])
    except Exception as e:
        print(name, "failed:", e)
