"""Stand-alone timing of the stride-2 3x3 layers of the W48 plan on both tcgen05 engines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ctypes as C
import torch
from rtpe_b200 import _lib as L
from _convutil import make_desc, pack_weights

SHAPES = [(64, 160, 160, 48, 96), (64, 160, 160, 48, 48), (64, 80, 80, 96, 192), (64, 80, 80, 48, 192),
          (64, 160, 160, 256, 96), (64, 320, 320, 64, 64), (64, 80, 80, 96, 96), (64, 80, 80, 48, 48)]


def run(lib, eng, n, h, w, cin, cout, reps=10):
    d, taps = make_desc(L.DT_BF16, eng, n, h, w, cin, cout, 3, 2, True)
    used = lib.brtpe_conv_select_engine(C.byref(d))
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, h, w, cin), generator=g).cuda().to(torch.bfloat16)
    wgt = (torch.randn((cout, cin, 3, 3), generator=g) / (cin * 9) ** 0.5).cuda()
    bias = torch.zeros(cout, device="cuda")
    out = torch.empty((n, h // 2, w // 2, d.out_ld), dtype=torch.bfloat16, device="cuda")
    packed = pack_weights(lib, wgt, taps, 3, d, used, True)
    plan = lib.brtpe_plan_create()
    L.check(lib.brtpe_plan_add_conv(plan, C.byref(d), L.ptr(x), L.ptr(packed), L.ptr(bias), None, L.ptr(out)), "add")
    st = L.stream_ptr()
    for _ in range(3):
        lib.brtpe_plan_run(plan, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.brtpe_plan_run(plan, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * n * (h // 2) * (w // 2) * 9 * cin * cout
    lib.brtpe_plan_destroy(plan)
    return used, ms, fl / ms / 1e9, out


def main():
    lib = L.load()
    for s in SHAPES:
        res = {}
        for eng in (L.ENGINE_UMMA, L.ENGINE_UMMA_HALO):
            used, ms, tf, out = run(lib, eng, *s)
            res[eng] = (ms, tf, out)
        same = (res[L.ENGINE_UMMA][2].float() - res[L.ENGINE_UMMA_HALO][2].float()).abs().max().item()
        print("s2 %3d->%3d @%dx%d x%d: per-tap %.4f ms %.0f TF | halo %.4f ms %.0f TF | max diff %.3g" %
              (s[3], s[4], s[1] // 2, s[2] // 2, s[0], res[L.ENGINE_UMMA][0], res[L.ENGINE_UMMA][1],
               res[L.ENGINE_UMMA_HALO][0], res[L.ENGINE_UMMA_HALO][1], same))


if __name__ == "__main__":
    main()
