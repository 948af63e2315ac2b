"""One conv shape through the halo kernel (pair or single mode by BRTPE_HALO_CG) vs the fp64 torch
reference; prints the error and the time.  usage: halo_pair_check.py N H W CIN COUT RES"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from rtpe_b200 import _lib as L  # noqa: E402
from _convutil import run_conv  # noqa: E402

n, h, w, cin, cout, res = [int(v) for v in sys.argv[1:7]]
got, ref, eng = run_conv(L.ENGINE_UMMA_HALO, "bf16", n, h, w, cin, cout, 3, 1, True, bool(res))
err = ((got - ref).abs().max() / ref.abs().max()).item()
bad = (~torch.isfinite(got)).sum().item()
print("shape", sys.argv[1:7], "engine", eng, "err %.3e" % err, "nonfinite", bad, "OK" if err <= 6e-3 and bad == 0 else "FAIL")
