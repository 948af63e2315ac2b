"""GPU: the CUDA path against the fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  Decode: bit-exact.  Model: fp32 mode <= 1e-4, bf16 <= 2e-2 of
the tensor max."""
import glob
import os

import numpy as np
import pytest
import torch

import rtpe_b200
from oracle.weights import fill_params_deterministic
from test_golden import GOLD, load_decode_fixture, split_people

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "decode_*.npz"))))
def test_decode_matches_reference_fixture(cuda_device, path):
    z, det, tag, tpj = load_decode_fixture(path)
    hp = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False, tag_per_joint=tpj)
    tk = hp.top_k(torch.from_numpy(det).cuda(), torch.from_numpy(tag).cuda())
    assert np.array_equal(tk["val_k"], z["val_k"])
    live = z["val_k"] > 0
    assert np.array_equal(tk["loc_k"][live], z["loc_k"].astype(np.int64)[live])
    assert np.array_equal(tk["tag_k"][live], z["tag_k"][live])
    got = hp.parse_batch(torch.from_numpy(det).cuda(), torch.from_numpy(tag).cuda(), True, True)
    want = split_people(z)
    for (gp, gs), (wp, ws) in zip(got, want):
        gp = np.asarray(gp, np.float32).reshape((-1,) + wp.shape[1:])
        assert np.array_equal(gp, wp)
        assert np.array_equal(np.asarray(gs, np.float32), ws)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_hhrnet_matches_reference_fixture(cuda_device, mode, tol):
    z = np.load(os.path.join(GOLD, "hhrnet_64x96.npz"))
    net = rtpe_b200.PoseHigherResolutionNet()
    fill_params_deterministic(net, int(z["seed"]))
    net = net.eval()
    model = net.cuda() if mode == "fp32" else rtpe_b200.network_to_half(net).cuda().eval()
    with torch.no_grad():
        y0, y1 = model(torch.from_numpy(z["x"]).cuda())
    for got, want in ((y0, z["y0"]), (y1, z["y1"])):
        want = torch.from_numpy(want)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert ((got.cpu() - want).abs().max() / want.abs().max()).item() <= tol
