"""GPU parity of the HeatmapParser drop-in against the CPU oracle (oracle/group_ref.py,
itself pinned to the reference's group.py).  Everything here is BIT-EXACT: NMS values,
top-k values / indices / tags, person count, person order, coordinates, scores."""
import numpy as np
import pytest
import torch

import rtpe_b200
from oracle import group_ref as G

pytestmark = pytest.mark.gpu

PARSER_KW = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
                 use_detection_val=True, ignore_too_much=False)


def make_parser(**over):
    kw = dict(PARSER_KW)
    extra = {k: over.pop(k) for k in list(over) if k in ("tag_per_joint", "nms_ksize", "nms_padding")}
    kw.update(over)
    return rtpe_b200.HeatmapParser(**kw, **extra), G.DecodeParams(**kw, **extra)


def assert_people_equal(got, want):
    assert len(got) == len(want)
    for (gp, gs), (wp, ws) in zip(got, want):
        wp = np.asarray(wp)
        assert gp.shape == wp.shape, (gp.shape, wp.shape)
        assert np.array_equal(gp, wp)
        assert len(gs) == len(ws)
        assert np.array_equal(np.asarray(gs, np.float32), np.asarray(ws, np.float32))


@pytest.mark.parametrize("h,w,k", [(64, 64, 5), (96, 128, 5), (50, 70, 3), (33, 47, 5), (64, 200, 1)])
def test_nms_exact(cuda_device, h, w, k):
    det = torch.randn(2, 17, h, w, generator=torch.Generator().manual_seed(h * w))
    det[0, 0, :8, :8] = 0.5                      # plateau: every equal maximum survives
    det[1, 3] = -det[1, 3].abs()                 # negative peaks stay negative
    hp, _ = make_parser(nms_ksize=k, nms_padding=(k - 1) // 2)
    got = hp.nms(det.cuda()).cpu().numpy()
    want = G.nms_ref(det.numpy(), k, (k - 1) // 2)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("h,w,t,tpj", [(64, 64, 1, True), (96, 128, 2, True), (61, 83, 1, True),
                                       (128, 96, 1, False), (320, 320, 1, True)])
def test_topk_exact(cuda_device, h, w, t, tpj):
    det, tag = rtpe_b200.synth_decode_batch(3, height=h, width=w, tag_dims=t, tag_per_joint=tpj,
                                            seed=11)
    hp, p = make_parser(tag_per_joint=tpj)
    got = hp.top_k(det.cuda(), tag.cuda())
    want = G.top_k_ref(det.numpy(), tag.numpy(), p)
    for key in ("val_k", "loc_k", "tag_k"):
        assert got[key].dtype == want[key].dtype, key
        assert np.array_equal(got[key], want[key]), key


def test_topk_sparse_and_negative_maps(cuda_device):
    """fewer than K positive peaks: zero-valued positions in index order, then negative
    peaks -- the canonical order of topk over the NMS'd map."""
    n, j, h, w = 2, 17, 40, 56
    det = torch.zeros(n, j, h, w)
    det[0, 0, 5, 7] = 0.9
    det[0, 0, 20, 30] = 0.4
    det[0, 1] = -1.0                              # all-negative plateau: every pixel a peak
    det[0, 2] = -torch.rand(h, w, generator=torch.Generator().manual_seed(3)) - 0.1
    det[1, 4, 0, 0] = 0.3
    tag = torch.randn(n, j, h, w, 1, generator=torch.Generator().manual_seed(4))
    hp, p = make_parser()
    got = hp.top_k(det.cuda(), tag.cuda())
    want = G.top_k_ref(det.numpy(), tag.numpy(), p)
    for key in ("val_k", "loc_k", "tag_k"):
        assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("seed,h,w,t,people", [(0, 96, 128, 1, 8), (1, 128, 128, 2, 20),
                                               (2, 320, 320, 1, 30), (3, 75, 101, 1, 12)])
def test_parse_batch_exact(cuda_device, seed, h, w, t, people):
    det, tag = rtpe_b200.synth_decode_batch(4, height=h, width=w, tag_dims=t, max_people=people,
                                            seed=100 + seed)
    hp, p = make_parser()
    got = hp.parse_batch(det.cuda(), tag.cuda(), True, True)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), p, True, True)
    assert_people_equal(got, want)


@pytest.mark.parametrize("adjust,refine", [(False, False), (True, False), (False, True)])
def test_parse_flags(cuda_device, adjust, refine):
    det, tag = rtpe_b200.synth_decode_batch(2, height=80, width=96, max_people=6, seed=77)
    hp, p = make_parser()
    got = hp.parse_batch(det.cuda(), tag.cuda(), adjust, refine)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), p, adjust, refine)
    assert_people_equal(got, want)


def test_parse_reference_structure(cuda_device):
    """N=1 parse returns the reference's structure: ([ (P,J,3+T) f32 ], [np.float32...])."""
    det, tag = rtpe_b200.synth_decode_batch(1, height=64, width=64, max_people=5, seed=5)
    hp, p = make_parser()
    ans, scores = hp.parse(det.cuda(), tag.cuda(), True, True)
    wp, ws = G.parse_image_ref(det.numpy().copy(), tag.numpy().copy(), p, True, True)
    assert isinstance(ans, list) and len(ans) == 1
    assert ans[0].dtype == np.float32 and np.array_equal(ans[0], wp)
    assert all(isinstance(s, np.float32) for s in scores)
    assert np.array_equal(np.asarray(scores), np.asarray(ws))


def test_adversarial_ties_and_collisions(cuda_device):
    """tags quantised to bf16 (duplicate dict keys), persons closer than the tag threshold
    (assignment ties), more than 30 peaks per joint."""
    det, tag = rtpe_b200.synth_decode_batch(6, height=96, width=96, max_people=30, seed=900)
    tag = tag.to(torch.bfloat16).to(torch.float32)
    tag = (tag * 0.25)                           # people 0.375 apart: many candidates per person
    hp, p = make_parser()
    got = hp.parse_batch(det.cuda(), tag.cuda(), True, True)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), p, True, True)
    assert_people_equal(got, want)


def test_more_than_30_people(cuda_device):
    """the person list is not capped at max_num_people."""
    h = w = 160
    det = torch.rand(1, 17, h, w, generator=torch.Generator().manual_seed(1)) * 0.01
    tag = torch.zeros(1, 17, h, w, 1)
    pid = 0
    for y in range(10, 150, 20):
        for x in range(10, 150, 20):
            if pid >= 40:
                break
            for j in range(17):
                yy, xx = y + (j % 4), x + (j // 4)
                det[0, j, yy, xx] = 0.5 + 0.01 * ((pid + 7 * j) % 40) + 0.0001 * j
                tag[0, j, yy - 2:yy + 3, xx - 2:xx + 3, 0] = 3.0 * pid
            pid += 1
    hp, p = make_parser()
    got = hp.parse_batch(det.cuda(), tag.cuda(), True, True)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), p, True, True)
    assert want[0][0].shape[0] > 30
    assert_people_equal(got, want)


def test_empty_detections(cuda_device):
    det = torch.rand(2, 17, 48, 48, generator=torch.Generator().manual_seed(2)) * 0.05
    tag = torch.zeros(2, 17, 48, 48, 1)
    hp, p = make_parser()
    got = hp.parse_batch(det.cuda(), tag.cuda(), True, True)
    for people, scores in got:
        assert people.size == 0 and scores == []
    ans, scores = hp.parse(det.cuda(), tag.cuda(), True, True)
    assert len(ans) == 1 and ans[0].size == 0 and scores == []


def test_match_adjust_refine_methods(cuda_device):
    """the individual public methods keep the reference's signatures and results."""
    det, tag = rtpe_b200.synth_decode_batch(2, height=72, width=88, max_people=7, seed=31)
    hp, p = make_parser()
    tk = G.top_k_ref(det.numpy(), tag.numpy(), p)
    got = hp.match(tk["tag_k"], tk["loc_k"], tk["val_k"])
    want = G.match_ref(tk["tag_k"], tk["loc_k"], tk["val_k"], p)
    for g, w_ in zip(got, want):
        assert np.array_equal(g, w_)
    got = hp.adjust(got, det)
    want = G.adjust_ref(want, det.numpy())
    for g, w_ in zip(got, want):
        assert np.array_equal(g, w_)
    for i in range(min(3, len(want[0]))):
        a = hp.refine(det[0].numpy(), tag[0].numpy(), got[0][i].copy())
        b = G.refine_ref(det[0].numpy(), tag[0].numpy(), want[0][i].copy())
        assert np.array_equal(a, b)


def test_errors_are_loud(cuda_device):
    hp, _ = make_parser(max_num_people=100)
    det, tag = rtpe_b200.synth_decode_batch(1, height=32, width=32)
    with pytest.raises(rtpe_b200.BrtpeError):
        hp.top_k(det.cuda(), tag.cuda())
    hp2 = rtpe_b200.HeatmapParser(nms_ksize=5, nms_padding=1, **PARSER_KW)
    with pytest.raises(rtpe_b200.BrtpeError):
        hp2.nms(det.cuda())


@pytest.mark.parametrize("k,people,dense", [(3, 30, True), (7, 30, True), (5, 40, True), (5, 64, False),
                                            (1, 30, True)])
def test_topk_stream_kernel_variants(cuda_device, k, people, dense):
    """the bulk-copy streaming top-k (rows 16-byte aligned) for every NMS radius, for K > 32 (two
    list slots per lane) and for dense maps where almost every row segment holds a candidate."""
    g = torch.Generator().manual_seed(31 * k + people)
    if dense:
        det = torch.randn(3, 17, 72, 256, generator=g) * 0.2
        tag = torch.randn(3, 17, 72, 256, 2, generator=g)
    else:
        det, tag = rtpe_b200.synth_decode_batch(3, height=200, width=132, tag_dims=2, seed=9)
    hp, p = make_parser(max_num_people=people, nms_ksize=k, nms_padding=(k - 1) // 2)
    got = hp.top_k(det.cuda(), tag.cuda())
    want = G.top_k_ref(det.numpy(), tag.numpy(), p)
    for key in ("val_k", "loc_k", "tag_k"):
        assert np.array_equal(got[key], want[key]), key


def test_topk_stream_full_size_bands(cuda_device):
    """640 x 640 maps of one image: 17 planes split into many row bands per plane (the global
    threshold word and the merge kernel), against the oracle."""
    det, tag = rtpe_b200.synth_decode_batch(1, height=640, width=640, tag_dims=2, seed=21)
    det = det + 0.02 * torch.randn(det.shape, generator=torch.Generator().manual_seed(3))
    hp, p = make_parser()
    got = hp.top_k(det.cuda(), tag.cuda())
    want = G.top_k_ref(det.numpy(), tag.numpy(), p)
    for key in ("val_k", "loc_k", "tag_k"):
        assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("t", [1, 3])
def test_refine_more_missing_persons_than_one_pass(cuda_device, t):
    """40 persons that all miss joints 0-2 (their peaks stay below the detection threshold): the
    streaming refine kernel evaluates them in two passes of 32 persons over the same maps."""
    h = w = 160
    g = torch.Generator().manual_seed(4)
    det = torch.rand(1, 17, h, w, generator=g) * 0.01
    tag = torch.randn(1, 17, h, w, t, generator=g) * 0.01
    pid = 0
    for y in range(10, 150, 20):
        for x in range(10, 150, 20):
            if pid >= 40:
                break
            for j in range(17):
                yy, xx = y + (j % 4), x + (j // 4)
                amp = 0.05 if j < 3 else 0.5                 # joints 0-2: never detected
                det[0, j, yy, xx] = amp + 0.0005 * ((pid + 7 * j) % 40) + 0.00001 * j
                tag[0, j, yy - 2:yy + 3, xx - 2:xx + 3, :] = 3.0 * pid
            pid += 1
    hp, p = make_parser()
    got = hp.parse_batch(det.cuda(), tag.cuda(), True, True)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), p, True, True)
    assert want[0][0].shape[0] >= 40
    assert (np.asarray(want[0][0])[:, :3, 2] > 0).any()      # refine filled missing joints
    assert_people_equal(got, want)


def test_person_capacity_overflow_retry(cuda_device):
    """more persons than the optimistic person capacity: the decode is enqueued for the small
    capacity, the flag read at the end triggers ONE retry at the J*K bound, same result."""
    h = w = 160
    det = torch.rand(2, 17, h, w, generator=torch.Generator().manual_seed(1)) * 0.01
    tag = torch.zeros(2, 17, h, w, 1)
    pid = 0
    for y in range(10, 150, 20):
        for x in range(10, 150, 20):
            if pid >= 24:
                break
            for j in range(17):
                yy, xx = y + (j % 4), x + (j // 4)
                det[0, j, yy, xx] = 0.5 + 0.01 * ((pid + 7 * j) % 40) + 0.0001 * j
                tag[0, j, yy - 2:yy + 3, xx - 2:xx + 3, 0] = 3.0 * pid
            pid += 1
    det[1], tag[1] = det[0].flip(-1), tag[0].flip(-2)
    hp, p = make_parser()
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), p, True, True)
    assert want[0][0].shape[0] >= 24
    small = rtpe_b200.HeatmapParser(**PARSER_KW, person_capacity=8)
    got = small.parse_batch(det.cuda(), tag.cuda(), True, True)
    assert_people_equal(got, want)
    assert_people_equal(hp.parse_batch(det.cuda(), tag.cuda(), True, True), want)
    # the overflow is remembered: the next decode starts at the J*K bound (no second grouping pass)
    assert small._capacity_hint == small._pmax_full() and hp._capacity_hint == 0
    val_k, ind_k, _, tag_k = small.top_k_device(det.cuda(), tag.cuda())
    ans, count, pmax, flag = small.match_device(val_k, ind_k, tag_k, w, defer_overflow=True)
    assert pmax == small._pmax_full() and flag is None
    assert_people_equal(small.parse_batch(det.cuda(), tag.cuda(), True, True), want)


def test_refine_ties_scalar_tags(cuda_device):
    """scalar tags + quantised heat maps: thousands of pixels tie in det - rint(|tag - mean|), the winner is
    the first index (np.argmax).  Guards the chunk-level tag-range bound of the streaming refine kernel
    (T == 1): a chunk may only be skipped when its bound is strictly below the person's best so far."""
    h, w = 96, 160
    g = torch.Generator().manual_seed(11)
    det = (torch.randint(0, 3, (2, 17, h, w), generator=g).float() * 0.0125)          # {0, .0125, .025}
    tag = torch.randint(0, 4, (2, 17, h, w, 1), generator=g).float() * 2.0            # {0, 2, 4, 6}
    for n in range(2):
        for pid in range(5):
            for j in range(17):
                if (j + pid) % 3 == 0:
                    continue                                  # joints this person misses
                y, x = 8 + 16 * pid + (j % 5), 10 + 8 * j
                det[n, j, y, x] = 0.5 + 0.01 * pid
                tag[n, j, y - 1:y + 2, x - 1:x + 2, 0] = 2.0 * pid
    hp, p = make_parser()
    got = hp.parse_batch(det.cuda(), tag.cuda(), True, True)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), p, True, True)
    assert_people_equal(got, want)
