"""Per-role cycle breakdown of conv_halo_kernel (brtpe_debug_halo_prof counters).
usage: halo_prof.py [N H W C] ...   (default: the five dominant W48 shapes at N = 16)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from rtpe_b200 import _lib as L  # noqa: E402
from _convutil import make_desc, pack_weights  # noqa: E402

NAMES = ["kernel", "prologue", "prod.loop", "prod.wait_freeA", "prod.wait_freeB", "mma.loop",
         "mma.wait_A", "mma.wait_B", "mma.wait_acc", "epi.loop", "epi.wait_acc", "tiles"]


def run(lib, n, h, w, c, reps=5):
    use_res = os.environ.get("RES", "1") != "0"
    d, taps = make_desc(L.DT_BF16, 0, n, h, w, c, c, 3, 1, True, res_ld=c if use_res else 0)
    used = lib.brtpe_conv_select_engine(C.byref(d))
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, h, w, c), generator=g).cuda().to(torch.bfloat16)
    wgt = (torch.randn((c, c, 3, 3), generator=g) / (c * 9) ** 0.5).cuda()
    bias = torch.zeros(c, device="cuda")
    res = torch.randn((n, h, w, c), generator=g).cuda().to(torch.bfloat16)
    out = torch.empty((n, h, w, d.out_ld), dtype=torch.bfloat16, device="cuda")
    packed = pack_weights(lib, wgt, taps, 3, d, used, True)
    plan = lib.brtpe_plan_create()
    L.check(lib.brtpe_plan_add_conv(plan, C.byref(d), L.ptr(x), L.ptr(packed), L.ptr(bias),
                                    L.ptr(res) if use_res else None, L.ptr(out)), "add")
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
    buf = torch.zeros((1024, 16), dtype=torch.int64, device="cuda")
    lib.brtpe_debug_halo_prof(L.ptr(buf), 1024)
    lib.brtpe_plan_run(plan, st)
    torch.cuda.synchronize()
    lib.brtpe_debug_halo_prof(None, 0)
    b = buf.cpu().double()
    live = b[:, 0] > 0
    b = b[live]
    fl = 2.0 * n * h * w * 9 * c * c
    print("N=%d %dx%d C=%d: %.4f ms %.1f TF, %d CTAs, tiles/CTA mean %.2f max %d" %
          (n, h, w, c, ms, fl / ms / 1e9, b.shape[0], b[:, 11].mean(), int(b[:, 11].max())))
    kt = b[:, 0].mean()
    print("   kernel cycles: mean %.0f  min %.0f  max %.0f" % (kt, b[:, 0].min(), b[:, 0].max()))
    for i in range(1, 11):
        print("   %-16s %8.0f  %5.1f%% of kernel" % (NAMES[i], b[:, i].mean(), 100 * b[:, i].mean() / kt))
    lib.brtpe_plan_destroy(plan)


def main():
    lib = L.load()
    a = [int(v) for v in sys.argv[1:]]
    shapes = [a[i:i + 4] for i in range(0, len(a), 4)] or [
        [16, 160, 160, 48], [16, 80, 80, 96], [16, 40, 40, 192], [16, 20, 20, 384], [16, 320, 320, 48],
        [32, 160, 160, 48], [32, 80, 80, 96]]
    for s in shapes:
        run(lib, *s)


if __name__ == "__main__":
    main()
