import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from rtpe_b200 import _lib as L
from _convutil import run_conv
def err(got, ref): return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()
shape = tuple(int(v) for v in sys.argv[1:8]) + (True, False)
got, ref, used = run_conv(L.ENGINE_UMMA_HALO, "bf16", *shape)
print(shape, dict((k, v) for k, v in os.environ.items() if k.startswith("BRTPE_")), "used", used, "err %.2e" % err(got, ref))
