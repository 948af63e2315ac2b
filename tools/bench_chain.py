"""Chained BasicBlock launch (brtpe_conv_chain_run's kernel through a plan) against the two single launches:
device time per block, same box.  usage: bench_chain.py N H W C [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from rtpe_b200 import _lib as L  # noqa: E402
from _convutil import make_desc, pack_weights  # noqa: E402


def main():
    n, h, w, c = [int(v) for v in sys.argv[1:5]]
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    lib = L.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn((n, h, w, c), generator=g).to(dev).to(torch.bfloat16)
    ws = [(torch.randn((c, c, 3, 3), generator=g) / (c * 9) ** 0.5).to(dev) for _ in range(2)]
    bs = [(torch.randn((c,), generator=g) * 0.1).to(dev) for _ in range(2)]
    d0, taps = make_desc(L.DT_BF16, L.ENGINE_UMMA_HALO, n, h, w, c, c, 3, 1, True)
    d1, _ = make_desc(L.DT_BF16, L.ENGINE_UMMA_HALO, n, h, w, c, c, 3, 1, True, res_ld=c)
    pk = [pack_weights(lib, ws[i], taps, 3, d0, L.ENGINE_UMMA_HALO, True) for i in range(2)]
    mid = torch.empty((n, h, w, d0.out_ld), dtype=torch.bfloat16, device=dev)
    out = torch.empty((n, h, w, d1.out_ld), dtype=torch.bfloat16, device=dev)
    flops = 2 * 2.0 * n * h * w * c * c * 9
    for mode in ("0", "2"):
        os.environ["BRTPE_CHAIN"] = mode
        plan = lib.brtpe_plan_create()
        nblocks = 4                               # x -> out -> x ... like a branch of four BasicBlocks
        src, dst = x, out
        for b in range(nblocks):
            L.check(lib.brtpe_plan_add_conv(plan, C.byref(d0), L.ptr(src), L.ptr(pk[0]), L.ptr(bs[0]), None,
                                            L.ptr(mid)), "add")
            L.check(lib.brtpe_plan_add_conv(plan, C.byref(d1), L.ptr(mid), L.ptr(pk[1]), L.ptr(bs[1]),
                                            L.ptr(src), L.ptr(dst)), "add")
            src, dst = dst, src
        st = L.stream_ptr()
        for _ in range(3):
            L.check(lib.brtpe_plan_graph_launch(plan, st), "launch")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            L.check(lib.brtpe_plan_graph_launch(plan, st), "launch")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps / nblocks
        print("N=%d %dx%d C=%d BRTPE_CHAIN=%s: %.4f ms per block (2 convs), %.1f TFLOP/s"
              % (n, h, w, c, mode, ms, flops / ms / 1e9))
        lib.brtpe_plan_destroy(plan)


if __name__ == "__main__":
    main()
