"""Sweep the halo conv kernel over cluster sizes / batch sizes for the W48 3x3 shapes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from rtpe_b200 import _lib as L  # noqa: E402
from _convutil import make_desc, pack_weights  # noqa: E402


def time_conv(lib, eng, n, h, w, cin, cout, k, stride, reps=20, flush=None):
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
    lib.brtpe_plan_destroy(plan)
    return ms, fl / ms / 1e9, used


def main():
    lib = L.load()
    shapes = [(160, 48), (80, 96), (40, 192), (20, 384), (320, 48), (160, 64)]
    for n in (8, 16, 32):
        for hw, c in shapes:
            row = []
            for cs in (1, 2, 4):
                os.environ["BRTPE_HALO_CS"] = str(cs)
                ms, tf, used = time_conv(lib, 0, n, hw, hw, c, c, 3, 1)
                row.append("cs%d %.4f ms %6.1f TF" % (cs, ms, tf))
            print("N=%-3d %3dx%-3d C=%-3d | %s" % (n, hw, hw, c, " | ".join(row)), flush=True)
    os.environ.pop("BRTPE_HALO_CS", None)
    for n in (8, 16):
        for (h, cin, cout, k, s) in [(160, 64, 256, 1, 1), (160, 256, 64, 1, 1), (160, 48, 96, 3, 2),
                                     (80, 96, 48, 1, 1), (40, 192, 48, 1, 1), (20, 384, 48, 1, 1)]:
            ms, tf, used = time_conv(lib, 0, n, h, h, cin, cout, k, s)
            print("N=%-3d %3dx%-3d %d->%d k%d s%d engine %d: %.4f ms %6.1f TF" % (n, h, h, cin, cout, k, s, used, ms, tf), flush=True)


if __name__ == "__main__":
    main()
