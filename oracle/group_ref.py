"""ORACLE (test infrastructure only) -- CPU restatement of the reference's
associative-embedding decode, ``rtpe/third_party/group.py`` (all line numbers
below are into that file under /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product path
(realtime-pose-estimation_b200/) never does.

Pinning: the reference has no tests or golden vectors for this path.  The
restatement is pinned against the reference's own ``HeatmapParser`` imported
in-process from /root/reference (tests/test_oracle_vs_reference.py, run when the
reference tree is present) and against tests/golden/*.npz that were produced by
the reference itself (oracle/make_golden.py).  The Hungarian step comes from the
un-vendored PyPI package ``munkres`` -> oracle/munkres_ref.py (PARITY UNPINNED
for that step, see its header).

Canonical top-k tie rule (the reference's is implementation-defined,
group.py:153-154): descending value, ascending flat index among equal values,
with -0.0 == +0.0.

All arrays are numpy; ``det`` is (N, J, H, W) float32, ``tag`` is
(N, Jt, H, W, T) float32 with Jt = J (tag_per_joint) or 1.
"""
from __future__ import annotations

import numpy as np

from .munkres_ref import Munkres


class DecodeParams:
    """Mirror of Params + HeatmapParser ctor arguments (group.py:100-132)."""

    def __init__(self, num_joints=17, max_num_people=30, detection_threshold=0.1,
                 tag_threshold=1.0, use_detection_val=True, ignore_too_much=False,
                 tag_per_joint=True, nms_ksize=5, nms_padding=2,
                 munkres_start_rule="previous"):
        self.num_joints = num_joints
        self.max_num_people = max_num_people
        self.detection_threshold = detection_threshold
        self.tag_threshold = tag_threshold
        self.use_detection_val = use_detection_val
        self.ignore_too_much = ignore_too_much
        self.tag_per_joint = tag_per_joint
        self.nms_ksize = nms_ksize
        self.nms_padding = nms_padding
        self.munkres_start_rule = munkres_start_rule


# ---------------------------------------------------------------------------
# NMS + top-k  (group.py:134-138, :144-179)
# ---------------------------------------------------------------------------
def nms_ref(det: np.ndarray, ksize: int, padding: int) -> np.ndarray:
    """det * (maxpool_{k,1,p}(det) == det); pool padding is -inf."""
    n, j, h, w = det.shape
    ho = h + 2 * padding - ksize + 1
    wo = w + 2 * padding - ksize + 1
    assert ho == h and wo == w, "reference NMS needs a shape-preserving pool"
    padded = np.full((n, j, h + 2 * padding, w + 2 * padding), -np.inf, np.float32)
    padded[:, :, padding:padding + h, padding:padding + w] = det
    mx = np.full_like(det, -np.inf)
    for dy in range(ksize):
        for dx in range(ksize):
            np.maximum(mx, padded[:, :, dy:dy + h, dx:dx + w], out=mx)
    return det * (mx == det).astype(np.float32)


def top_k_ref(det: np.ndarray, tag: np.ndarray, p: DecodeParams):
    """-> dict(tag_k (N,J,K,T) f32, loc_k (N,J,K,2) i64 [x,y], val_k (N,J,K) f32)."""
    n, j, h, w = det.shape
    k = p.max_num_people
    flat = nms_ref(det, p.nms_ksize, p.nms_padding).reshape(n, j, h * w)
    # canonical order: value descending, index ascending among equals.
    # stable argsort of -value does exactly that (and -0.0 == +0.0 compare equal).
    order = np.argsort(-flat, axis=2, kind="stable")[:, :, :k]
    val_k = np.take_along_axis(flat, order, axis=2)
    t = tag.reshape(tag.shape[0], tag.shape[1], h * w, -1)
    if not p.tag_per_joint:
        t = np.broadcast_to(t, (t.shape[0], j, h * w, t.shape[3]))
    tag_k = np.stack([np.take_along_axis(t[:, :, :, i], order, axis=2)
                      for i in range(t.shape[3])], axis=3)
    x = order % w
    y = order // w                       # == (ind / w).long() for all probed sizes
    loc_k = np.stack((x, y), axis=3).astype(np.int64)
    return {"tag_k": np.ascontiguousarray(tag_k, np.float32),
            "loc_k": loc_k, "val_k": np.ascontiguousarray(val_k, np.float32)}


# ---------------------------------------------------------------------------
# grouping  (group.py:26-97)
# ---------------------------------------------------------------------------
def match_by_tag_ref(tag_k, loc_k, val_k, p: DecodeParams) -> np.ndarray:
    """One image.  tag_k (J,K,T) f32, loc_k (J,K,2) i64, val_k (J,K) f32 ->
    (P, J, 3+T) float32, persons in creation order (empty -> shape (0,))."""
    nj = p.num_joints
    width = 3 + tag_k.shape[2]
    persons = {}        # key (np.float32 tag[0]) -> (J, 3+T) float64
    taglists = {}       # key -> list of (T,) float32 tag vectors
    for idx in range(nj):                                   # joint_order = 0..J-1
        cand = np.concatenate((loc_k[idx].astype(np.float64),
                               val_k[idx][:, None].astype(np.float64),
                               tag_k[idx].astype(np.float64)), axis=1)
        keep = cand[:, 2] > p.detection_threshold
        cand = cand[keep]
        ctag = tag_k[idx][keep]
        if cand.shape[0] == 0:
            continue
        if idx == 0 or len(persons) == 0:
            for tvec, row in zip(ctag, cand):
                key = tvec[0]
                if key not in persons:
                    persons[key] = np.zeros((nj, width))
                persons[key][idx] = row
                taglists[key] = [tvec]
            continue
        keys = list(persons.keys())[:p.max_num_people]
        means = np.array([np.mean(taglists[kk], axis=0) for kk in keys])
        if p.ignore_too_much and len(keys) == p.max_num_people:
            continue
        diff = cand[:, None, 3:] - means[None, :, :]
        dist = np.linalg.norm(diff, ord=2, axis=2)
        dist_saved = dist.copy()
        cost = dist
        if p.use_detection_val:
            cost = np.round(dist) * 100 - cand[:, 2:3]
        n_add, n_grp = dist.shape
        if n_add > n_grp:
            cost = np.concatenate((cost, np.zeros((n_add, n_add - n_grp)) + 1e10), axis=1)
        pairs = Munkres(p.munkres_start_rule).compute(cost)
        for r, c in pairs:
            if r < n_add and c < n_grp and dist_saved[r][c] < p.tag_threshold:
                key = keys[c]
                persons[key][idx] = cand[r]
                taglists[key].append(ctag[r])
            else:
                key = ctag[r][0]
                if key not in persons:
                    persons[key] = np.zeros((nj, width))
                persons[key][idx] = cand[r]
                taglists[key] = [ctag[r]]
    return np.array([persons[kk] for kk in persons]).astype(np.float32)


def match_ref(tag_k, loc_k, val_k, p: DecodeParams):
    return [match_by_tag_ref(t, l, v, p) for t, l, v in zip(tag_k, loc_k, val_k)]


# ---------------------------------------------------------------------------
# adjust  (group.py:181-200)
# ---------------------------------------------------------------------------
def adjust_ref(ans, det: np.ndarray):
    """Quarter-pixel shift towards the larger neighbour, then +0.5; in place."""
    h, w = det.shape[2], det.shape[3]
    for b, people in enumerate(ans):
        if people.size == 0:
            continue
        for pi in range(people.shape[0]):
            for ji in range(people.shape[1]):
                if people[pi, ji, 2] > 0:
                    col = people[pi, ji, 0]          # x
                    row = people[pi, ji, 1]          # y
                    r, c = int(row), int(col)
                    plane = det[b, ji]
                    if plane[r, min(c + 1, w - 1)] > plane[r, max(c - 1, 0)]:
                        col = col + np.float32(0.25)
                    else:
                        col = col - np.float32(0.25)
                    if plane[min(r + 1, h - 1), c] > plane[max(0, r - 1), c]:
                        row = row + np.float32(0.25)
                    else:
                        row = row - np.float32(0.25)
                    people[pi, ji, 0] = col + np.float32(0.5)
                    people[pi, ji, 1] = row + np.float32(0.5)
    return ans


# ---------------------------------------------------------------------------
# refine  (group.py:202-264)
# ---------------------------------------------------------------------------
def refine_ref(det: np.ndarray, tag: np.ndarray, keypoints: np.ndarray):
    """det (J,H,W) f32, tag (J,H,W,T) f32, keypoints (J,3+T) f32 (modified)."""
    if tag.ndim == 3:
        tag = tag[:, :, :, None]
    nj, h, w = det.shape
    seen = []
    for i in range(nj):
        if keypoints[i, 2] > 0:
            x, y = keypoints[i][:2].astype(np.int32)
            seen.append(tag[i, y, x])
    mean_tag = np.mean(seen, axis=0)
    found = np.zeros((nj, 3))
    for i in range(nj):
        plane = det[i]
        dist = ((tag[i] - mean_tag[None, None, :]) ** 2).sum(axis=2) ** 0.5
        score = plane - np.round(dist)
        y, x = np.unravel_index(np.argmax(score), plane.shape)
        val = plane[y, x]
        fx = x + 0.5
        fy = y + 0.5
        if plane[y, min(x + 1, w - 1)] > plane[y, max(x - 1, 0)]:
            fx += 0.25
        else:
            fx -= 0.25
        if plane[min(y + 1, h - 1), x] > plane[max(0, y - 1), x]:
            fy += 0.25
        else:
            fy -= 0.25
        found[i] = (fx, fy, val)
    for i in range(nj):
        if found[i, 2] > 0 and keypoints[i, 2] == 0:
            keypoints[i, :2] = found[i, :2]
            keypoints[i, 2] = found[i, 2]
    return keypoints


# ---------------------------------------------------------------------------
# parse  (group.py:266-287)
# ---------------------------------------------------------------------------
def parse_image_ref(det1: np.ndarray, tag1: np.ndarray, p: DecodeParams,
                    adjust=True, refine=True):
    """Reference ``parse`` applied to ONE image (det1 (1,J,H,W), tag1 (1,Jt,H,W,T)).
    Returns (people, scores): people = (P,J,3+T) f32 array (or shape-(0,) array),
    scores = list of np.float32."""
    tk = top_k_ref(det1, tag1, p)
    ans = match_ref(tk["tag_k"], tk["loc_k"], tk["val_k"], p)
    if adjust:
        ans = adjust_ref(ans, det1)
    scores = [person[:, 2].mean() for person in ans[0]]
    people = ans[0]
    if refine:
        det_np = det1[0]
        tag_np = tag1[0]
        if not p.tag_per_joint:
            tag_np = np.tile(tag_np, (p.num_joints, 1, 1, 1))
        for i in range(len(people)):
            people[i] = refine_ref(det_np, tag_np, people[i])
    return people, scores


def parse_batch_ref(det: np.ndarray, tag: np.ndarray, p: DecodeParams,
                    adjust=True, refine=True):
    """The batched drop-in's definition: reference parse per image."""
    return [parse_image_ref(det[i:i + 1], tag[i:i + 1], p, adjust, refine)
            for i in range(det.shape[0])]


# ---------------------------------------------------------------------------
# explicit reduction orders -- what the CUDA kernels
# implement; checked against numpy itself in tests/test_oracle_group.py
# ---------------------------------------------------------------------------
def pairwise_sum_f32(vals) -> np.float32:
    """numpy's float32 add.reduce order for a 1-D run of n < 128 values."""
    a = [np.float32(v) for v in vals]
    n = len(a)
    if n < 8:
        s = np.float32(0.0)
        for v in a:
            s = np.float32(s + v)
        return s
    r = a[:8]
    i = 8
    while i < n - (n % 8):
        r = [np.float32(r[q] + a[i + q]) for q in range(8)]
        i += 8
    s = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3]))
                   + np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
    while i < n:
        s = np.float32(s + a[i])
        i += 1
    return s


def mean_tags_f32(taglist) -> np.ndarray:
    """np.mean(list of (T,) f32 vectors, axis=0): T == 1 -> pairwise rule along
    the list; T >= 2 -> sequential row-by-row accumulation; then / n in f32."""
    arr = np.asarray(taglist, np.float32)
    n, t = arr.shape
    out = np.zeros((t,), np.float32)
    if t == 1:
        out[0] = pairwise_sum_f32(arr[:, 0])
    else:
        for c in range(t):
            s = arr[0, c]
            for r in range(1, n):
                s = np.float32(s + arr[r, c])
            out[c] = s
    return (out / np.float32(n)).astype(np.float32)
