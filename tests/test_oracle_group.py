"""CPU: properties of the decode oracle and of the explicit float32 reduction orders the
CUDA kernels implement, checked against numpy itself."""
import numpy as np
import pytest
import torch

import rtpe_b200
from oracle import group_ref as G


def test_pairwise_sum_matches_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(1, 34)) + [40, 64, 100, 127]:
        for _ in range(50):
            a = rng.standard_normal(n).astype(np.float32)
            strided = np.zeros((n, 4), np.float32)
            strided[:, 2] = a
            assert G.pairwise_sum_f32(a) == strided[:, 2].sum()
            assert G.pairwise_sum_f32(a) == a.sum()


@pytest.mark.parametrize("t", [1, 2, 3])
def test_mean_tags_matches_numpy(t):
    rng = np.random.default_rng(1)
    for n in range(1, 19):
        for _ in range(40):
            lst = [rng.standard_normal(t).astype(np.float32) for _ in range(n)]
            assert np.array_equal(G.mean_tags_f32(lst), np.mean(lst, axis=0))


def test_nms_matches_torch_maxpool():
    det = torch.randn(2, 5, 37, 41, generator=torch.Generator().manual_seed(0))
    det[0, 0, :6, :6] = 1.5
    for k in (1, 3, 5, 7):
        pool = torch.nn.MaxPool2d(k, 1, (k - 1) // 2)
        want = det * torch.eq(pool(det), det).float()
        assert np.array_equal(G.nms_ref(det.numpy(), k, (k - 1) // 2), want.numpy())


def test_topk_matches_torch_on_distinct_values():
    det, tag = rtpe_b200.synth_decode_batch(2, height=48, width=64, tag_dims=2, seed=3)
    p = G.DecodeParams()
    got = G.top_k_ref(det.numpy(), tag.numpy(), p)
    pool = torch.nn.MaxPool2d(5, 1, 2)
    nd = det * torch.eq(pool(det), det).float()
    val, ind = nd.view(2, 17, -1).topk(30, dim=2)
    pos = val.numpy() > 0
    assert np.array_equal(got["val_k"], val.numpy())
    assert np.array_equal(got["loc_k"][..., 0][pos], (ind % 64).numpy()[pos])
    assert np.array_equal(got["loc_k"][..., 1][pos], (ind // 64).numpy()[pos])


def test_person_list_is_not_capped_and_keys_collapse():
    k, j = 30, 17
    tag_k = np.zeros((j, k, 1), np.float32)
    loc_k = np.zeros((j, k, 2), np.int64)
    val_k = np.zeros((j, k), np.float32)
    # joint 0: 30 candidates, two of them share tag[0] -> 29 persons
    tag_k[0, :, 0] = np.arange(k) * 5.0
    tag_k[0, 7, 0] = tag_k[0, 3, 0]
    val_k[0] = 0.9
    # joint 1: 30 candidates far from everyone -> all start new persons (list grows past 30)
    tag_k[1, :, 0] = 1000.0 + np.arange(k) * 5.0
    val_k[1] = 0.8
    out = G.match_by_tag_ref(tag_k, loc_k, val_k, G.DecodeParams())
    assert out.shape == (29 + 30, j, 4)
    assert out.dtype == np.float32


def test_empty_image_returns_empty_array():
    z = np.zeros((17, 30), np.float32)
    out = G.match_by_tag_ref(np.zeros((17, 30, 1), np.float32), np.zeros((17, 30, 2), np.int64), z,
                             G.DecodeParams())
    assert out.shape == (0,)


def test_adjust_and_refine_fill_missing_joint():
    det, tag = rtpe_b200.synth_decode_batch(1, height=64, width=64, max_people=3, seed=9)
    p = G.DecodeParams()
    people, scores = G.parse_image_ref(det.numpy().copy(), tag.numpy().copy(), p, True, False)
    assert people.shape[0] >= 1
    frac = people[..., :2] - np.floor(people[..., :2])
    live = people[..., 2] > 0
    assert np.all(np.isin(frac[live], (0.25, 0.75)))
    kp = people[0].copy()
    kp[5] = 0
    out = G.refine_ref(det[0].numpy(), tag[0].numpy(), kp.copy())
    assert out[5, 2] > 0
