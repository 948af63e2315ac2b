"""CPU: the C-ABI library loads here (no GPU) and exports every symbol include/brtpe.h
declares; argument validation works without touching a device."""
import ctypes as C
import os
import re

from rtpe_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "brtpe.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(brtpe_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(L.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(L.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(L.exported_symbols()) == set(names)


def test_argument_validation_without_gpu():
    lib = L.load(require_cuda=False)
    assert lib.brtpe_version() >= 100
    rc = lib.brtpe_nms(None, None, 1, 8, 8, 5, 2, None)
    assert rc == -1 and b"brtpe_nms" in lib.brtpe_last_error()
    rc = lib.brtpe_nms_topk_gather(None, None, 1, 17, 17, 8, 8, 1, 30, 5, 2, None, None, None, None,
                                   None, 0, None)
    assert rc == -1
    d = L.ConvDesc()
    assert lib.brtpe_conv_select_engine(C.byref(d)) < 0
    assert lib.brtpe_topk_workspace_bytes(1, 17, 64, 64, 30) > 0
    assert lib.brtpe_group_workspace_bytes(2, 17, 30, 1, 64) > 0
    assert lib.brtpe_refine_workspace_bytes(2, 17, 1, 64) > 0


def test_product_fails_loudly_without_cuda():
    import pytest
    import torch
    import rtpe_b200
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    hp = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False)
    with pytest.raises(rtpe_b200.BrtpeError):
        hp.parse(torch.zeros(1, 17, 32, 32), torch.zeros(1, 17, 32, 32, 1))
    with pytest.raises(rtpe_b200.BrtpeError):
        rtpe_b200.PoseHigherResolutionNet().eval()(torch.zeros(1, 3, 64, 64))
