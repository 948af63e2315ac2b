"""The remaining BASELINE.json configurations on the GPU path.

* config 3 -- multi-scale (2.0 / 1.0 / 0.5) + flip-test aggregation + AE grouping: the product
  pipeline against the oracle pipeline (oracle forward in fp32 on the CPU, oracle aggregation,
  oracle decode).  Network outputs are compared within the bf16 tolerance of BASELINE.json
  (max|delta|/max|ref| <= 2e-2), the aggregation of the DEVICE outputs within 1e-5, and the decode
  of the device maps is bit-exact.
* config 5 -- decode-only stress at the full size: batch 1024, 17 joints, 320x320, K = 30.  The
  oracle needs ~0.3 s per image, so it checks a sample of the batch bit-exactly; the whole batch
  is covered by size-independent properties (shards of the batch decode to the same rows,
  a permuted batch decodes to the permuted result).
"""
import numpy as np
import pytest
import torch

import rtpe_b200
from rtpe_b200 import inference
from oracle import group_ref as G
from oracle.aggregate_ref import aggregate_flip_multiscale_ref
from oracle.hhrnet_ref import hhrnet_forward_ref
from oracle.weights import fill_params_deterministic

pytestmark = pytest.mark.gpu

PARSER_KW = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
                 use_detection_val=True, ignore_too_much=False)


def _rel(got, ref):
    return ((got.double().cpu() - ref.double()).abs().max() / ref.double().abs().max()).item()


def test_config3_multiscale_flip_pipeline(cuda_device):
    net = rtpe_b200.PoseHigherResolutionNet()
    fill_params_deterministic(net, 11)
    net.eval()
    sd = net.state_dict()
    base_h, base_w = 64, 96
    g = torch.Generator().manual_seed(12)
    # descending scales like legacy/valid_ae_avg.py:166; H and W stay multiples of 32
    xs = [(2.0, torch.randn(2, 3, 2 * base_h, 2 * base_w, generator=g)),
          (1.0, torch.randn(2, 3, base_h, base_w, generator=g)),
          (0.5, torch.randn(2, 3, base_h // 2 + 0, base_w // 2 + 16, generator=g))]
    model = rtpe_b200.network_to_half(net).cuda().eval()
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)

    dev_outs, ref_outs = [], []
    with torch.no_grad():
        for scale, x in xs:
            xf = torch.flip(x, [3])
            y0, y1 = model(torch.cat((x, xf), 0).cuda())
            ref = hhrnet_forward_ref(sd, x)
            ref_f = hhrnet_forward_ref(sd, xf)
            n = x.shape[0]
            for got, want in ((y0[:n], ref[0]), (y1[:n], ref[1]), (y0[n:], ref_f[0]), (y1[n:], ref_f[1])):
                assert _rel(got, want) <= 2e-2
            dev_outs.append((scale, [y0[:n].cpu(), y1[:n].cpu()], [y0[n:].cpu(), y1[n:].cpu()]))
            ref_outs.append((scale, ref, ref_f))
    det, tag = pipe.forward_aggregate_multiscale([(s, x.cuda()) for s, x in xs], (base_h, base_w))
    rdet, rtag = aggregate_flip_multiscale_ref(dev_outs, (base_w, base_h))
    assert det.shape == rdet.shape == (2, 17, base_h, base_w)
    assert tag.shape == rtag.shape == (2, 17, base_h, base_w, 2)
    assert (det.cpu() - rdet).abs().max() <= 1e-5 * rdet.abs().max()
    assert (tag.cpu() - rtag).abs().max() <= 1e-5 * rtag.abs().max()
    # end to end against the all-fp32 oracle pipeline: bf16 budget
    odet, otag = aggregate_flip_multiscale_ref(ref_outs, (base_w, base_h))
    assert _rel(det, odet) <= 2e-2 and _rel(tag, otag) <= 2e-2
    # decode of the device maps: bit-exact
    got = parser.parse_batch(det, tag, True, True)
    want = G.parse_batch_ref(det.cpu().numpy().copy(), tag.cpu().numpy().copy(),
                             G.DecodeParams(**PARSER_KW), True, True)
    for (gp, gs), (wp, ws) in zip(got, want):
        assert gp.shape == np.asarray(wp).shape and np.array_equal(gp, wp)
        assert np.array_equal(np.asarray(gs, np.float32), np.asarray(ws, np.float32))


def test_config5_decode_stress_full_size(cuda_device):
    n_total, h, w = 1024, 320, 320
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    chunks = []
    for i0 in range(0, n_total, 128):       # generated in shards: image i only depends on (seed, i)
        chunks.append(rtpe_b200.synth_decode_batch(128, height=h, width=w, seed=1234, device="cuda",
                                                   first_index=i0))
    det = torch.cat([c[0] for c in chunks], 0)
    tag = torch.cat([c[1] for c in chunks], 0)
    del chunks
    assert det.shape == (n_total, 17, h, w) and tag.shape == (n_total, 17, h, w, 1)
    ans, count, scores = parser.decode_device(det, tag, True, True)
    torch.cuda.synchronize()
    count_h = count.cpu()
    assert int(count_h.min()) >= 1 and int(count_h.max()) <= 17 * 30

    # (1) a sample of the batch against the oracle, bit-exact
    sample = [0, 1, 2, 3, 255, 256, 511, 777, 1022, 1023]
    want = G.parse_batch_ref(det[sample].cpu().numpy().copy(), tag[sample].cpu().numpy().copy(),
                             G.DecodeParams(**PARSER_KW), True, True)
    ans_h, sc_h = ans[sample].cpu().numpy(), scores[sample].cpu().numpy()
    for k, i in enumerate(sample):
        c = int(count_h[i])
        wp, ws = want[k]
        wp = np.asarray(wp)
        assert wp.shape == (c, 17, 4)
        assert np.array_equal(ans_h[k, :c], wp)
        assert np.array_equal(sc_h[k, :c], np.asarray(ws, np.float32))

    # (2) images are independent: a shard decodes to the same rows as inside the full batch
    a2, c2, s2 = parser.decode_device(det[256:512].contiguous(), tag[256:512].contiguous(), True, True)
    assert torch.equal(c2, count[256:512])
    pm = min(a2.shape[1], ans.shape[1])
    live = (torch.arange(pm, device=det.device)[None, :] < c2[:, None])
    assert torch.equal(a2[:, :pm][live], ans[256:512, :pm][live])
    assert torch.equal(s2[:, :pm][live], scores[256:512, :pm][live])

    # (3) a permuted batch decodes to the permuted result
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(3)).cuda()
    a3, c3, s3 = parser.decode_device(det[:64][perm].contiguous(), tag[:64][perm].contiguous(), True, True)
    assert torch.equal(c3, count[:64][perm])
    pm = min(a3.shape[1], ans.shape[1])
    live = (torch.arange(pm, device=det.device)[None, :] < c3[:, None])
    assert torch.equal(a3[:, :pm][live], ans[:64][perm][:, :pm][live])


@pytest.mark.parametrize("early", [None, 0, 1, 3])
def test_run_stream_equals_run_device(cuda_device, early):
    """the pipelined host loop (double-buffered H2D copies, part of the next copy released while the
    network runs) returns, batch by batch, exactly what the device-resident call returns -- different
    host batches in flight must never mix."""
    net = rtpe_b200.PoseHigherResolutionNet()
    fill_params_deterministic(net, 5)
    model = rtpe_b200.network_to_half(net).cuda().eval()
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)
    g = torch.Generator().manual_seed(2)
    batches = [torch.randn(3, 3, 64, 96, generator=g).pin_memory() for _ in range(4)]
    want = []
    for xb in batches:
        ans, count, scores = pipe.run_device(xb.cuda())
        want.append((ans.cpu(), count.cpu(), scores.cpu()))
    got = [(a.cpu(), c.cpu(), s.cpu()) for a, c, s in pipe.run_stream(iter(batches), early_images=early)]
    assert len(got) == len(want)
    for (ga, gc, gs), (wa, wc, ws) in zip(got, want):
        assert torch.equal(gc, wc)
        for i in range(ga.shape[0]):
            n = int(gc[i])
            assert torch.equal(ga[i, :n], wa[i, :n]) and torch.equal(gs[i, :n], ws[i, :n])
    assert list(pipe.run_stream(iter([]))) == []
