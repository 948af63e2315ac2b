"""GPU parity of PoseHigherResolutionNet.forward against the CPU oracle
(oracle/hhrnet_ref.py, pinned to the reference module).  Metric per output tensor:
max|delta| / max|ref|.  fp32 mode <= 1e-4, bf16 mode <= 2e-2 (BASELINE.json north_star)."""
import pytest
import torch

import rtpe_b200
from rtpe_b200 import _lib as L
from oracle.hhrnet_ref import hhrnet_forward_ref

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4
TOL_BF16 = 2e-2


def _rel(got, ref):
    return ((got.double().cpu() - ref.double()).abs().max() / ref.double().abs().max()).item()


@pytest.fixture(scope="module")
def model_and_ref():
    torch.manual_seed(0)
    net = rtpe_b200.PoseHigherResolutionNet().eval()
    # non-trivial BN statistics so that the folding is actually exercised
    g = torch.Generator().manual_seed(1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    x = torch.randn(3, 3, 128, 192, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = hhrnet_forward_ref(net.state_dict(), x)
    return net, x, ref


def test_forward_fp32(cuda_device, model_and_ref):
    net, x, ref = model_and_ref
    net = net.float().cuda()
    net.chunk_size = 2                    # 3 images -> chunks of 2 + 1
    with torch.no_grad():
        out = net(x.cuda())
    assert [tuple(o.shape) for o in out] == [tuple(r.shape) for r in ref]
    for o, r in zip(out, ref):
        assert o.dtype == torch.float32
        assert _rel(o, r) <= TOL_FP32


@pytest.mark.parametrize("engine", [L.ENGINE_FFMA, L.ENGINE_AUTO])
def test_forward_bf16(cuda_device, model_and_ref, engine):
    net, x, ref = model_and_ref
    import copy
    half = rtpe_b200.network_to_half(copy.deepcopy(net).float()).cuda().eval()
    half[1].conv_engine = engine
    with torch.no_grad():
        out = half(x.cuda())
    for o, r in zip(out, ref):
        assert o.dtype == torch.float32
        assert torch.isfinite(o).all()
        assert _rel(o, r) <= TOL_BF16


def test_graph_and_eager_agree(cuda_device, model_and_ref):
    net, x, ref = model_and_ref
    net = net.float().cuda()
    with torch.no_grad():
        net.use_cuda_graph = True
        a = net(x.cuda())
        a2 = net(x.cuda())
        net.use_cuda_graph = False
        net.invalidate_plans()
        b = net(x.cuda())
    for u, v, w_ in zip(a, a2, b):
        assert torch.equal(u, v) and torch.equal(u, w_)


def test_requires_cuda_and_eval(cuda_device):
    net = rtpe_b200.PoseHigherResolutionNet()
    with pytest.raises(rtpe_b200.BrtpeError):
        net.eval()(torch.zeros(1, 3, 64, 64))
    with pytest.raises(rtpe_b200.BrtpeError):
        net.train().cuda()(torch.zeros(1, 3, 64, 64).cuda())
    with pytest.raises(ValueError):
        net.eval().cuda()(torch.zeros(1, 3, 70, 64).cuda())


@pytest.mark.parametrize("chunk", [8, 3])
def test_forward_flip_pair_equals_materialised_batch(cuda_device, chunk):
    """the fused flip-test batch (the stem's im2col reads the mirrored half straight from x) gives the
    outputs of forward(cat(x, flip(x))): identical for a float32 model input, and for the
    network_to_half wrapper up to the fp16 output rounding the wrapper's half outputs carry."""
    import rtpe_b200
    from oracle.weights import fill_params_deterministic
    net = rtpe_b200.PoseHigherResolutionNet()
    fill_params_deterministic(net, 9)
    model = rtpe_b200.network_to_half(net).cuda().eval()
    model[1].chunk_size = chunk
    x = torch.randn(5, 3, 64, 96, generator=torch.Generator().manual_seed(3)).cuda()
    both = torch.cat((x, torch.flip(x, [3])), 0)
    with torch.no_grad():
        # (a) the half-precision network fed float32 directly: bit-identical
        ya = model[1](both)
        yb = model[1].forward_flip_pair(x)
        for a, b in zip(ya, yb):
            assert a.shape == b.shape and torch.equal(a, b)
        # (b) the wrapper (input rounded through fp16, outputs through fp16 in the wrapper only)
        yc = model(both)
        yd = model[1].forward_flip_pair(x, via_half=True)
        for c, d in zip(yc, yd):
            assert c.shape == d.shape
            assert (c - d.float()).abs().max().item() <= 1e-3 * c.abs().max().item()


def test_fuse_in_epilogue_matches_fuse_kernel(cuda_device, model_and_ref):
    """HRNet cross-resolution fuse-add folded into the conv epilogues (opt-in) against the stand-alone
    fuse_sum path and the oracle: same network, sums taken in a different order / with bf16 partial
    sums at low resolution, so equal within the bf16 budget, and the plan has no fuse op left."""
    net, x, ref = model_and_ref
    import copy
    half = rtpe_b200.network_to_half(copy.deepcopy(net).float()).cuda().eval()
    with torch.no_grad():
        half[1].fuse_in_epilogue = False
        a = [o.clone() for o in half(x.cuda())]
        half[1].fuse_in_epilogue = True
        b = [o.clone() for o in half(x.cuda())]
    plans = [p for k, p in half[1]._plans.items() if k[-1] is True]
    assert plans and all(op[0] != "fuse" for op in plans[0].recorder.ops)
    assert any(op[0] == "conv" and op[1].n_add > 0 for op in plans[0].recorder.ops)
    for u, v, r in zip(a, b, ref):
        assert _rel(v, r) <= TOL_BF16 and _rel(u, r) <= TOL_BF16
        assert _rel(u, v.cpu()) <= 2 * TOL_BF16        # two results, each within the budget of the oracle


@pytest.mark.parametrize("name", ["two_deconvs_k3_heads", "bottleneck_stages_k3_k2_deconvs",
                                  "no_deconv_shared_tag"])
@pytest.mark.parametrize("mode,tol", [("fp32", TOL_FP32), ("bf16", TOL_BF16)])
def test_constructor_variants_vs_oracle(cuda_device, name, mode, tol):
    """BOTTLENECK stages, 3x3 final convs, two deconv stages with kernel sizes 4 / 3 / 2 and per-stage
    concat flags (pose_higher_hrnet.py:266-287, :447-546) against the oracle (pinned to the reference
    for the same configurations in tests/test_oracle_vs_reference.py)."""
    from oracle.variants import variant_kwargs
    from oracle.weights import fill_params_deterministic
    net = rtpe_b200.PoseHigherResolutionNet(**variant_kwargs(name))
    fill_params_deterministic(net, 31)
    net.eval()
    x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(32))
    with torch.no_grad():
        ref = hhrnet_forward_ref(net.state_dict(), x)
    model = net.cuda() if mode == "fp32" else rtpe_b200.network_to_half(net).cuda().eval()
    with torch.no_grad():
        out = model(x.cuda())
    assert [tuple(o.shape) for o in out] == [tuple(r.shape) for r in ref]
    for o, r in zip(out, ref):
        assert _rel(o, r) <= tol


def test_chained_blocks_agree(cuda_device, model_and_ref, monkeypatch):
    """the plan's chain pass (conv1 -> conv2 of every BasicBlock of the 48-channel branches as one
    L2-blocked launch, plan.cu fuse_chains) changes the schedule, not one bit of the result; graph and
    eager replay, programmatic dependent launch on and off"""
    net, x, ref = model_and_ref
    import copy
    half = rtpe_b200.network_to_half(copy.deepcopy(net).float()).cuda().eval()
    half[1].chunk_size = 4
    xs = torch.cat([x, x.flip(3)], 0)[:4].cuda()
    outs = {}
    with torch.no_grad():
        for chain in ("0", "2"):
            for graph in (True, False):
                monkeypatch.setenv("BRTPE_CHAIN", chain)
                half[1].use_cuda_graph = graph
                half[1].invalidate_plans()
                outs[(chain, graph)] = [o.clone() for o in half(xs)]
                again = half(xs)
                for u, v in zip(outs[(chain, graph)], again):
                    assert torch.equal(u, v)
    base = outs[("0", True)]
    for key, val in outs.items():
        for u, v in zip(base, val):
            assert torch.equal(u, v), key
    for o, r in zip(base, ref):
        assert _rel(o[:3], r) <= TOL_BF16
