"""In-process ablation of conv_halo_kernel: BRTPE_HALO_DBG variants x shapes (one python start)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from rtpe_b200 import _lib as L  # noqa: E402
from _convutil import make_desc, pack_weights  # noqa: E402


def bench(lib, n, h, w, c, dbg, reps=30):
    os.environ["BRTPE_HALO_DBG"] = str(dbg)
    d, taps = make_desc(L.DT_BF16, 0, n, h, w, c, c, 3, 1, True, res_ld=c)
    used = lib.brtpe_conv_select_engine(C.byref(d))
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, h, w, c), generator=g).cuda().to(torch.bfloat16)
    wgt = (torch.randn((c, c, 3, 3), generator=g) / (c * 9) ** 0.5).cuda()
    bias = torch.zeros(c, device="cuda")
    res = torch.randn((n, h, w, c), generator=g).cuda().to(torch.bfloat16)
    out = torch.empty((n, h, w, d.out_ld), dtype=torch.bfloat16, device="cuda")
    packed = pack_weights(lib, wgt, taps, 3, d, used, True)
    plan = lib.brtpe_plan_create()
    L.check(lib.brtpe_plan_add_conv(plan, C.byref(d), L.ptr(x), L.ptr(packed), L.ptr(bias), L.ptr(res),
                                    L.ptr(out)), "add")
    st = L.stream_ptr()
    for _ in range(3):
        lib.brtpe_plan_graph_launch(plan, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.brtpe_plan_graph_launch(plan, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    lib.brtpe_plan_destroy(plan)
    return ms


def main():
    lib = L.load()
    dbgs = [int(v) for v in os.environ.get("DBGS", "0,16,15,14,30,2,18").split(",")]
    shapes = [[16, 160, 160, 48], [16, 80, 80, 96], [16, 40, 40, 192], [16, 20, 20, 384], [32, 80, 80, 96],
              [32, 160, 160, 48]]
    print("%-22s" % "shape" + "".join("%10s" % ("dbg=%d" % d) for d in dbgs))
    for n, h, w, c in shapes:
        fl = 2.0 * n * h * w * 9 * c * c
        row = [bench(lib, n, h, w, c, d) for d in dbgs]
        print("%-22s" % ("%dx%dx%d C=%d" % (n, h, w, c)) + "".join("%10.4f" % m for m in row) +
              "   ms | TF(dbg=%d) %.0f" % (dbgs[0], fl / row[0] / 1e9))


if __name__ == "__main__":
    main()
