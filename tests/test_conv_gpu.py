"""GPU parity of the convolution engine against a float64 torch convolution of the same
(bf16-rounded) operands.  Metric: max|delta| / max|ref| per output tensor.
fp32 (FFMA) <= 1e-5; bf16 storage of the output bounds the bf16 path at 2^-8 relative."""
import pytest
import torch

from rtpe_b200 import _lib as L
from _convutil import run_conv, run_conv_fused, run_conv_split

pytestmark = pytest.mark.gpu

SHAPES = [
    # n, h, w, cin, cout, k, stride, relu, res
    (2, 32, 48, 48, 48, 3, 1, True, True),
    (1, 40, 40, 96, 96, 3, 1, True, False),
    (2, 20, 20, 192, 192, 3, 1, False, True),
    (1, 20, 20, 384, 384, 3, 1, True, True),
    (2, 32, 32, 64, 256, 1, 1, True, True),
    (1, 32, 32, 256, 64, 1, 1, True, False),
    (2, 32, 32, 48, 96, 3, 2, True, False),
    (1, 64, 64, 64, 64, 3, 2, True, False),
    (1, 16, 16, 256, 48, 3, 1, True, False),
    (3, 24, 40, 48, 34, 1, 1, False, False),
    (1, 16, 16, 384, 48, 1, 1, False, False),
    # Cout tilings that do not cover the channels exactly (176 = 2 x 96, 208 = 2 x 112): the
    # chunk-wise fast epilogue must not run over the end (AttentionStudentSteps, inplanes 80)
    (2, 32, 32, 176, 176, 1, 1, True, False),
    (1, 24, 24, 64, 208, 1, 1, True, True),
    (1, 16, 24, 320, 32, 1, 1, True, True),
]


def _err(got, ref):
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("shape", SHAPES)
def test_ffma_fp32(cuda_device, shape):
    got, ref, eng = run_conv(L.ENGINE_FFMA, "fp32", *shape)
    assert eng == L.ENGINE_FFMA
    assert _err(got, ref) <= 1e-5


THIN_SHAPES = [
    # Cout <= 16: conv_ffma_thin_kernel (256 pixels x 16 channels per CTA) in fp32 mode
    (2, 32, 40, 48, 12, 3, 1, True, False),
    (1, 24, 24, 51, 16, 3, 1, True, True),      # Cin not a multiple of 4: scalar loads
    (3, 17, 23, 48, 1, 3, 1, False, False),
    (2, 16, 16, 96, 8, 1, 1, True, False),
    (1, 32, 32, 64, 16, 3, 2, True, False),
]


@pytest.mark.parametrize("shape", THIN_SHAPES)
def test_ffma_fp32_thin_layers(cuda_device, shape):
    got, ref, eng = run_conv(L.ENGINE_FFMA, "fp32", *shape)
    assert eng == L.ENGINE_FFMA
    assert torch.isfinite(got).all()
    assert _err(got, ref) <= 1e-5


@pytest.mark.parametrize("shape", SHAPES)
def test_ffma_bf16(cuda_device, shape):
    got, ref, eng = run_conv(L.ENGINE_FFMA, "bf16", *shape)
    assert _err(got, ref) <= 6e-3


@pytest.mark.parametrize("shape", SHAPES)
def test_umma_bf16(cuda_device, shape):
    got, ref, eng = run_conv(L.ENGINE_UMMA, "bf16", *shape)
    assert eng == L.ENGINE_UMMA
    assert torch.isfinite(got).all()
    assert _err(got, ref) <= 6e-3


HALO_SHAPES = [
    # n, h, w, cin, cout, k, stride, relu, res  (3x3 stride 1 only)
    (2, 32, 48, 48, 48, 3, 1, True, True),
    (1, 40, 40, 96, 96, 3, 1, True, False),
    (3, 20, 20, 192, 192, 3, 1, False, True),
    (1, 20, 20, 384, 384, 3, 1, True, True),
    (1, 16, 16, 256, 48, 3, 1, True, False),
    (2, 24, 40, 64, 64, 3, 1, True, False),
    (1, 17 * 2, 22, 48, 48, 3, 1, True, True),
    (5, 16, 8, 96, 96, 3, 1, True, True),
    # widths outside the teacher's {48, 64, 96, 192, 384}: the students' 16-channel dilation
    # branches, the 152 -> 160-channel mid stem, and odd multiples of 16
    (2, 32, 32, 48, 16, 3, 1, True, False),
    (2, 32, 32, 256, 160, 3, 1, True, False),
    (2, 32, 32, 160, 48, 3, 1, True, False),
    (1, 24, 24, 64, 224, 3, 1, True, True),
    (1, 24, 24, 64, 256, 3, 1, True, True),
    (1, 24, 24, 80, 144, 3, 1, False, False),
    (2, 16, 16, 48, 1, 3, 1, False, False),
    # MultistageStudent (256 + 18 -> 288 / 320 stored channels) and AttentionStudentSteps
    # (16-channel space-to-depth image, 4 x 64 space-to-depth planes, 64 / 96 / 112 / 176 wide concats)
    (1, 16, 24, 320, 288, 3, 1, True, False),
    (1, 16, 24, 288, 288, 3, 1, True, False),
    (1, 16, 24, 288, 32, 3, 1, True, False),
    (1, 21, 35, 256, 288, 3, 1, True, False),
    (2, 32, 48, 16, 64, 3, 1, True, False),
    (2, 16, 24, 256, 48, 3, 1, True, False),
    (2, 16, 24, 64, 16, 3, 1, True, False),
    (2, 16, 24, 112, 32, 3, 1, True, False),
    (1, 12, 20, 176, 48, 3, 1, True, False),
    (1, 12, 20, 176, 16, 3, 1, False, False),
]


@pytest.mark.parametrize("shape", HALO_SHAPES)
def test_umma_halo_bf16(cuda_device, shape):
    got, ref, eng = run_conv(L.ENGINE_UMMA_HALO, "bf16", *shape)
    assert eng == L.ENGINE_UMMA_HALO
    assert torch.isfinite(got).all()
    assert _err(got, ref) <= 6e-3


@pytest.mark.parametrize("knob,value", [("BRTPE_HALO_TPS", "1"), ("BRTPE_HALO_TPS", "3"),
                                        ("BRTPE_HALO_NT", "2"), ("BRTPE_HALO_NT", "3"),
                                        ("BRTPE_HALO_NO_RESIDENT", "1")])
def test_umma_halo_tilings(cuda_device, knob, value, monkeypatch):
    """weight-stage size (taps per stage), Cout tiling and resident/streamed weights are
    tuning choices: every setting must give the same result"""
    monkeypatch.setenv(knob, value)
    for shape in [(2, 32, 32, 96, 96, 3, 1, True, True), (1, 48, 40, 48, 48, 3, 1, True, True),
                  (3, 20, 20, 192, 192, 3, 1, True, True), (1, 16, 24, 64, 64, 3, 1, True, False)]:
        got, ref, eng = run_conv(L.ENGINE_UMMA_HALO, "bf16", *shape)
        assert eng == L.ENGINE_UMMA_HALO
        assert _err(got, ref) <= 6e-3


def test_umma_halo_many_items(cuda_device):
    """more work items than SMs: every CTA walks several tile pairs (ring wrap-around,
    accumulator double buffering, cross-tile residual prefetch), odd tile count at the end"""
    for shape in [(7, 80, 72, 48, 48, 3, 1, True, True), (3, 80, 80, 96, 96, 3, 1, True, True),
                  (9, 40, 40, 192, 192, 3, 1, True, True), (5, 36, 20, 384, 384, 3, 1, True, True)]:
        got, ref, eng = run_conv(L.ENGINE_UMMA_HALO, "bf16", *shape)
        assert eng == L.ENGINE_UMMA_HALO
        assert torch.isfinite(got).all()
        assert _err(got, ref) <= 6e-3


# ---- fp32 mode on the tensor cores: split (hi, lo) bf16 activations, three MMAs per product
SPLIT_SHAPES = [
    # n, h, w, cin, cout, k, stride, relu, res  (Cout a multiple of 16: the split epilogue's contract)
    (2, 32, 48, 48, 48, 3, 1, True, True),
    (1, 40, 40, 96, 96, 3, 1, True, False),
    (3, 20, 20, 192, 192, 3, 1, False, True),
    (1, 20, 20, 384, 384, 3, 1, True, True),
    (1, 16, 16, 256, 48, 3, 1, True, False),
    (2, 24, 40, 64, 64, 3, 1, True, False),
    (2, 32, 32, 64, 256, 1, 1, True, True),
    (1, 32, 32, 256, 64, 1, 1, True, False),
    (2, 32, 32, 48, 96, 3, 2, True, False),
    (1, 64, 64, 64, 64, 3, 2, True, False),
    (2, 32, 32, 192, 384, 3, 2, True, False),
    (3, 24, 40, 48, 32, 1, 1, False, False),
    (1, 16, 16, 384, 48, 1, 1, False, False),
    (2, 32, 32, 32, 64, 1, 1, True, False),
]


@pytest.mark.parametrize("shape", SPLIT_SHAPES)
def test_split_fp32_on_tcgen05(cuda_device, shape):
    got, ref, eng = run_conv_split(L.ENGINE_AUTO, *shape)
    assert eng in (L.ENGINE_UMMA, L.ENGINE_UMMA_HALO)
    assert torch.isfinite(got).all()
    # hi*hi + lo*hi + hi*lo drops the lo*lo term (2^-16 relative) and the output is rounded to
    # 16 significant bits: 3e-5 of the tensor max is the contract the network test builds on
    assert _err(got, ref) <= 3e-5


def test_split_fp32_per_tap_engine_and_many_items(cuda_device):
    for shape in [(2, 32, 48, 48, 48, 3, 1, True, True), (1, 40, 40, 96, 96, 3, 1, True, True)]:
        got, ref, eng = run_conv_split(L.ENGINE_UMMA, *shape)
        assert eng == L.ENGINE_UMMA and _err(got, ref) <= 3e-5
    for shape in [(7, 80, 72, 48, 48, 3, 1, True, True), (3, 80, 80, 96, 96, 3, 1, True, True),
                  (9, 40, 40, 192, 192, 3, 1, True, True)]:
        got, ref, eng = run_conv_split(L.ENGINE_UMMA_HALO, *shape)
        assert eng == L.ENGINE_UMMA_HALO and _err(got, ref) <= 3e-5


# ---- 3x3 / stride 2 on the halo engine: four pixel-parity planes per tile (one input load per pixel)
HALO_S2_SHAPES = [
    # n, h, w, cin, cout, k, stride, relu, res   (h, w = INPUT size)
    (2, 64, 64, 48, 96, 3, 2, True, False),
    (2, 64, 96, 48, 48, 3, 2, False, False),
    (3, 80, 80, 96, 192, 3, 2, True, False),
    (1, 160, 160, 64, 64, 3, 2, True, False),
    (2, 64, 64, 256, 96, 3, 2, True, False),
    (2, 80, 80, 48, 192, 3, 2, False, True),
    (5, 32, 48, 192, 256, 3, 2, True, False),
    (1, 34 * 2, 22 * 2, 48, 48, 3, 2, True, True),      # ragged tiles: 34 x 22 output pixels
    (7, 160, 144, 48, 96, 3, 2, True, False),           # more work items than SMs
]


@pytest.mark.parametrize("shape", HALO_S2_SHAPES)
def test_umma_halo_stride2_bf16(cuda_device, shape):
    got, ref, eng = run_conv(L.ENGINE_UMMA_HALO, "bf16", *shape)
    assert eng == L.ENGINE_UMMA_HALO
    assert torch.isfinite(got).all()
    assert _err(got, ref) <= 6e-3


@pytest.mark.parametrize("shape", HALO_S2_SHAPES[:6])
def test_umma_halo_stride2_split_fp32(cuda_device, shape):
    got, ref, eng = run_conv_split(L.ENGINE_UMMA_HALO, *shape)
    assert eng == L.ENGINE_UMMA_HALO
    assert _err(got, ref) <= 3e-5


# ---- HRNet cross-resolution fuse-add in the conv epilogue (pose_higher_hrnet.py:245-254)
FUSED_CASES = [
    # engine, n, h, w, cin, cout, k, stride, relu, res, addend shifts, two outputs
    (L.ENGINE_UMMA_HALO, 2, 64, 64, 48, 48, 3, 1, True, True, (1, 2, 3), True),     # y_0 of a 4-branch module
    (L.ENGINE_UMMA_HALO, 3, 32, 48, 48, 48, 3, 1, True, True, (1,), True),          # stage 2
    (L.ENGINE_UMMA_HALO, 7, 80, 72, 48, 48, 3, 1, True, True, (1, 2), True),        # many items
    (L.ENGINE_UMMA_HALO, 2, 64, 64, 48, 96, 3, 2, True, True, (1, 2), False),       # y_1: stride-2 halo, TMA stores
    (L.ENGINE_UMMA_HALO, 2, 64, 96, 48, 96, 3, 2, True, True, (), False),
    (L.ENGINE_UMMA, 2, 64, 64, 48, 192, 3, 2, True, True, (0, 1), False),           # y_2: per-tap engine
    (L.ENGINE_UMMA, 2, 32, 32, 48, 384, 3, 2, True, True, (0, 0), False),           # y_3
    (L.ENGINE_UMMA, 3, 40, 40, 96, 192, 3, 2, True, True, (1,), False),
    (L.ENGINE_UMMA, 2, 32, 32, 96, 48, 1, 1, False, False, (0, 2), False),
]


@pytest.mark.parametrize("case", FUSED_CASES)
def test_fuse_addends_in_epilogue(cuda_device, case):
    eng_req, rest = case[0], case[1:]
    got, got2, ref, ref2, eng = run_conv_fused(eng_req, *rest)
    assert eng == eng_req
    assert torch.isfinite(got).all()
    assert _err(got, ref) <= 6e-3
    if ref2 is not None:
        assert torch.isfinite(got2).all()
        assert _err(got2, ref2) <= 6e-3


@pytest.mark.parametrize("shape", [(8, 32, 32, 48), (6, 48, 40, 48), (64, 16, 24, 48), (6, 80, 80, 48),
                                   (4, 40, 24, 32), (16, 160, 160, 48), (2, 21, 35, 48)])
def test_halo_chain_bit_identical(cuda_device, shape):
    """BasicBlock body as one chained launch (conv1 -> conv2 + residual, sub-batches interleaved, layer 1
    waiting on per-image completion counters) == the two single launches, bit for bit, for the
    intermediate tensor and the output; ragged tiles, one / many items per CTA, several sub-batches"""
    from _convutil import run_chain
    r = run_chain(*shape)
    assert torch.isfinite(r["two"][1].float()).all()
    assert torch.equal(r["chain"][0], r["two"][0])
    assert torch.equal(r["chain"][1], r["two"][1])


@pytest.mark.parametrize("g", ["1", "2", "4"])
def test_halo_chain_sub_batch_sizes(cuda_device, g, monkeypatch):
    """the sub-batch size is a tuning choice; repeated launches re-use the self-resetting counters"""
    from _convutil import run_chain
    monkeypatch.setenv("BRTPE_CHAIN_G", g)
    for rep in range(2):
        r = run_chain(8, 48, 48, 48, seed=rep)
        assert torch.equal(r["chain"][0], r["two"][0])
        assert torch.equal(r["chain"][1], r["two"][1])

