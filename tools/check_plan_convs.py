"""Diagnostic: every conv of a recorded bf16 plan, run stand-alone on random data through the
engine the plan chose and through the CUDA-core FFMA engine (same bf16 operands), compared.
usage: python tools/check_plan_convs.py steps80|steps48|multistage"""
import ctypes as C
import copy
import sys

import torch

sys.path.insert(0, ".")
from rtpe_b200 import _lib as L                                    # noqa: E402
from rtpe_b200.students import AttentionStudentSteps, MultistageStudent   # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "steps80"
    torch.manual_seed(53)                       # module default init; the convs see random data anyway
    if which.startswith("steps"):
        net = AttentionStudentSteps(None, "cpu", int(which[5:]), 17, 1, True)
    else:
        net = MultistageStudent(None, "cpu", half_precision=True)
    net = net.eval().cuda()
    lib = L.load()
    dev = torch.device("cuda")
    R, _ = net._record(2, 128, 128, "bf16", dev, False, False)
    seen = set()
    g = torch.Generator().manual_seed(0)
    worst = 0.0
    for op in R.ops:
        if op[0] != "conv":
            continue
        _, d, x, packed, bias, residual, out = op
        key = (d.N, d.Hin, d.Win, d.Cin, d.in_ld, d.in_coff, d.ntaps, tuple(d.tap_dy[:d.ntaps]),
               tuple(d.tap_dx[:d.ntaps]), d.Cout, d.Cout_store, d.out_ld, d.out_coff, d.res_ld, d.relu)
        if key in seen:
            continue
        seen.add(key)
        eng = lib.brtpe_conv_select_engine(C.byref(d))
        xin = torch.randn((d.N, d.Hin, d.Win, d.in_ld), generator=g).to(dev, torch.bfloat16)
        rin = torch.randn((d.N, d.Hout, d.Wout, d.res_ld), generator=g).to(dev, torch.bfloat16) \
            if residual is not None else None
        outs = []
        for use_ffma in (False, True):
            dd = copy.copy(d)
            w = packed
            if use_ffma:
                dd.engine = L.ENGINE_FFMA
                if eng != L.ENGINE_FFMA:
                    w = packed[:, :d.Cout, :d.Cin].float().permute(0, 2, 1).contiguous()
            o = torch.zeros((d.N, d.Hout, d.Wout, d.out_ld), dtype=torch.bfloat16, device=dev)
            L.check(lib.brtpe_conv_run(C.byref(dd), L.ptr(xin), L.ptr(w), L.ptr(bias), L.ptr(rin),
                                       L.ptr(o), L.stream_ptr()), "brtpe_conv_run")
            torch.cuda.synchronize()
            outs.append(o.float())
        a, b = outs
        err = ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
        worst = max(worst, err)
        flag = "  <-- MISMATCH" if not err <= 1e-2 else ""
        print("eng %d  %dx%dx%d Cin %d(ld %d,+%d) taps %d dy%s -> Cout %d/%d (ld %d,+%d) res %d relu %d : %.2e%s"
              % (eng, d.N, d.Hin, d.Win, d.Cin, d.in_ld, d.in_coff, d.ntaps,
                 sorted(set(d.tap_dy[:d.ntaps])), d.Cout, d.Cout_store, d.out_ld, d.out_coff, d.res_ld,
                 d.relu, err, flag))
    print("worst", worst)


if __name__ == "__main__":
    main()
