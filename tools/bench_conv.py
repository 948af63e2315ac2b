"""Time one conv shape through brtpe_conv_run (for ncu captures and quick A/B tests).
usage: bench_conv.py ENGINE N H W CIN COUT K STRIDE [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from rtpe_b200 import _lib as L  # noqa: E402
from _convutil import make_desc, pack_weights  # noqa: E402


def main():
    eng, n, h, w, cin, cout, k, stride = [int(v) for v in sys.argv[1:9]]
    reps = int(sys.argv[9]) if len(sys.argv) > 9 else 20
    lib = L.load()
    d, taps = make_desc(L.DT_BF16, eng, n, h, w, cin, cout, k, stride, True, res_ld=cout)
    used = lib.brtpe_conv_select_engine(C.byref(d))
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, h, w, cin), generator=g).cuda().to(torch.bfloat16)
    wgt = (torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).cuda()
    bias = torch.zeros(cout, device="cuda")
    ho, wo = h // stride, w // stride
    res = torch.randn((n, ho, wo, cout), generator=g).cuda().to(torch.bfloat16)
    out = torch.empty((n, ho, wo, d.out_ld), dtype=torch.bfloat16, device="cuda")
    packed = pack_weights(lib, wgt, taps, k, d, used, True)
    plan = lib.brtpe_plan_create()
    L.check(lib.brtpe_plan_add_conv(plan, C.byref(d), L.ptr(x), L.ptr(packed), L.ptr(bias), L.ptr(res),
                                    L.ptr(out)), "add")
    st = L.stream_ptr()
    for _ in range(3):
        L.check(lib.brtpe_plan_run(plan, st), "run")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.brtpe_plan_run(plan, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * n * ho * wo * len(taps) * cin * cout
    print("engine %d (used %d) %dx%dx%d %d->%d k%d s%d: %.4f ms  %.1f TFLOP/s" %
          (eng, used, n, h, w, cin, cout, k, stride, ms, fl / ms / 1e9))
    lib.brtpe_plan_destroy(plan)


if __name__ == "__main__":
    main()
