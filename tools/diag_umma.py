"""GPU diagnostic for the tcgen05 conv kernel: prints where errors are (by channel, by
pixel, borders) so that a blind fix is possible from the log alone."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from rtpe_b200 import _lib as L  # noqa: E402
from _convutil import run_conv  # noqa: E402

CASES = [
    ("1x1 64->64 16x8", (1, 8, 16, 64, 64, 1, 1, False, False)),
    ("1x1 64->64 32x32", (1, 32, 32, 64, 64, 1, 1, False, False)),
    ("1x1 48->48", (1, 16, 16, 48, 48, 1, 1, False, False)),
    ("1x1 128->64", (1, 16, 16, 128, 64, 1, 1, False, False)),
    ("3x3 64->64", (1, 16, 16, 64, 64, 3, 1, False, False)),
    ("3x3 48->48 res relu", (2, 32, 48, 48, 48, 3, 1, True, True)),
    ("3x3 s2 48->96", (2, 32, 32, 48, 96, 3, 2, True, False)),
    ("3x3 384->384", (1, 20, 20, 384, 384, 3, 1, True, True)),
    ("1x1 48->34", (3, 24, 40, 48, 34, 1, 1, False, False)),
]


def main():
    print("device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for name, shape in CASES:
        try:
            got, ref, eng = run_conv(L.ENGINE_UMMA, "bf16", *shape)
        except Exception as e:  # noqa: BLE001
            print("CASE %-24s EXCEPTION %s" % (name, e))
            traceback.print_exc()
            # a trap leaves the context dead: stop here
            return 1
        err = (got - ref).abs()
        scale = ref.abs().max().item()
        nan = (~torch.isfinite(got)).sum().item()
        e = torch.nan_to_num(err, nan=1e9)
        print("CASE %-24s eng=%d max_rel=%.3e nan=%d ref_max=%.3f" %
              (name, eng, e.max().item() / scale, nan, scale))
        if e.max().item() / scale > 1e-2:
            bych = e.amax(dim=(0, 2, 3))
            print("   err by out-channel:", ["%.2g" % v for v in bych[:48].tolist()])
            byrow = e.amax(dim=(0, 1, 3))
            print("   err by out-row    :", ["%.2g" % v for v in byrow[:40].tolist()])
            bycol = e.amax(dim=(0, 1, 2))
            print("   err by out-col    :", ["%.2g" % v for v in bycol[:48].tolist()])
            print("   got[0,:4,0,:4]", got[0, :4, 0, :4].tolist())
            print("   ref[0,:4,0,:4]", ref[0, :4, 0, :4].tolist())
            ratio = (got / ref.clamp_min(1e-6))[0, :4, 4, 4:8]
            print("   got/ref sample", ratio.tolist())
    return 0


if __name__ == "__main__":
    sys.exit(main())
