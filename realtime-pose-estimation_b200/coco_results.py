"""COCO keypoint result records and OKS evaluation for the decoded people.

* ``build_keypoint_results`` / ``write_keypoint_results``: the result list the reference's
  ``CocoDataset.evaluate`` builds and dumps (rtpe/third_party/COCODataset.py:160-275: one record per
  person with ``image_id``, ``category_id``, 17 x (x, y, score) ``keypoints``, ``score`` and the tight
  ``bbox`` of the keypoints; no OKS-NMS -- the reference keeps every detection, :205-214), without
  the ``pycocotools`` / ``json_tricks`` dependencies.  Pinned against the reference's own method
  bodies in tests/test_coco_results.py (extracted from the source file, this container only).
* ``oks`` / ``evaluate_keypoints``: the keypoint variant of the COCO evaluation protocol that the
  reference delegates to ``pycocotools.cocoeval.COCOeval`` (:277-293).  pycocotools is a third-party
  dependency that is neither vendored under /root/reference nor installed here: this is a
  restatement of the published protocol (OKS with the per-keypoint constants, greedy matching of
  score-sorted detections at OKS thresholds 0.50:0.05:0.95, at most 20 detections per image,
  101-point interpolated precision, area ranges all / medium / large) -- **parity unpinned**.

Host-side numpy; nothing here touches the GPU.
"""
from __future__ import annotations

import json
from collections import OrderedDict, defaultdict

import numpy as np

NUM_JOINTS = 17
# per-keypoint constants of the COCO keypoint task (sigma_i; k_i = 2 * sigma_i)
COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87,
                        .89, .89]) / 10.0
STATS_NAMES = ["AP", "Ap .5", "AP .75", "AP (M)", "AP (L)", "AR", "AR .5", "AR .75", "AR (M)", "AR (L)"]


def process_keypoints(keypoints):
    """COCODataset.py:135-148 (returns a copy; the float() round trip changes nothing)."""
    return np.array(keypoints, copy=True)


def build_keypoint_results(preds, scores, image_ids, dataset_with_center=False,
                           test_ignore_center=False, category_id=1):
    """COCODataset.py:176-221 + :243-275.  ``preds``: per image a list of (J, 3 + T) arrays
    (x, y, score, tag...), ``scores``: per image the person scores, ``image_ids``: per image the COCO
    image id (the reference derives it from the file name, ``int(file_name[-16:-4])``).
    -> list of result records in the reference's order (images in first-seen order)."""
    kpts = defaultdict(list)
    for idx, people in enumerate(preds):
        for k, kpt in enumerate(people):
            kpt = np.asarray(kpt)
            area = (np.max(kpt[:, 0]) - np.min(kpt[:, 0])) * (np.max(kpt[:, 1]) - np.min(kpt[:, 1]))
            kpt = process_keypoints(kpt)
            if dataset_with_center and not test_ignore_center:
                kpt = kpt[:-1]
            kpts[int(image_ids[idx])].append({"keypoints": kpt[:, 0:3], "score": scores[idx][k],
                                              "tags": kpt[:, 3], "image": int(image_ids[idx]),
                                              "area": area})
    results = []
    for people in kpts.values():                       # images in first-seen order, every detection kept
        kp = np.stack([p["keypoints"] for p in people]).astype(np.float64)      # (P, 17, 3)
        lo, hi = kp.min(axis=1), kp.max(axis=1)                                 # tight keypoint box
        flat = kp.reshape(len(people), NUM_JOINTS * 3)
        for k, p in enumerate(people):
            results.append({"image_id": p["image"], "category_id": category_id,
                            "keypoints": list(flat[k]), "score": p["score"],
                            "bbox": [lo[k, 0], lo[k, 1], hi[k, 0] - lo[k, 0], hi[k, 1] - lo[k, 1]]})
    return results


def write_keypoint_results(results, res_file):
    """COCODataset.py:223-241: ``json.dump(results, sort_keys=True, indent=4)`` (plain floats)."""
    def plain(o):
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        raise TypeError(type(o))
    with open(res_file, "w") as f:
        json.dump(results, f, sort_keys=True, indent=4, default=plain)


def oks(gt_keypoints, gt_area, dt_keypoints, gt_bbox=None, sigmas=COCO_SIGMAS):
    """Object keypoint similarity of one detection against one ground truth.  ``gt_keypoints`` (17, 3)
    with visibility flags, ``dt_keypoints`` (17, 3).  Ground truths without a labelled keypoint are
    scored by the distance to the doubled bounding box, as in the COCO protocol."""
    g = np.asarray(gt_keypoints, dtype=np.float64).reshape(-1, 3)
    d = np.asarray(dt_keypoints, dtype=np.float64).reshape(-1, 3)
    var = (np.asarray(sigmas, dtype=np.float64) * 2) ** 2
    vis = g[:, 2] > 0
    if vis.any():
        dx, dy = d[:, 0] - g[:, 0], d[:, 1] - g[:, 1]
    else:
        x, y, w, h = gt_bbox
        x0, x1, y0, y1 = x - w, x + 2 * w, y - h, y + 2 * h
        z = np.zeros(len(d))
        dx = np.maximum(z, x0 - d[:, 0]) + np.maximum(z, d[:, 0] - x1)
        dy = np.maximum(z, y0 - d[:, 1]) + np.maximum(z, d[:, 1] - y1)
    e = (dx ** 2 + dy ** 2) / var / (gt_area + np.spacing(1)) / 2
    if vis.any():
        e = e[vis]
    return float(np.sum(np.exp(-e)) / e.shape[0])


def evaluate_keypoints(ground_truth, results, max_dets=20):
    """COCO keypoint evaluation (restated protocol, see the module docstring).
    ``ground_truth``: list of annotation dicts with ``image_id``, ``keypoints`` (51 values),
    ``num_keypoints``, ``area``, ``bbox`` and optional ``iscrowd``; ``results``: the records of
    ``build_keypoint_results``.  -> OrderedDict over STATS_NAMES (-1 where undefined)."""
    thrs = np.linspace(0.5, 0.95, 10)
    recs = np.linspace(0.0, 1.0, 101)
    ranges = [("all", 0.0, 1e10), ("medium", 32.0 ** 2, 96.0 ** 2), ("large", 96.0 ** 2, 1e10)]
    gts, dts = defaultdict(list), defaultdict(list)
    for g in ground_truth:
        gts[g["image_id"]].append(g)
    for d in results:
        dts[d["image_id"]].append(d)
    image_ids = sorted(set(gts) | set(dts))
    precision = -np.ones((len(thrs), len(recs), len(ranges)))
    recall = -np.ones((len(thrs), len(ranges)))
    for a, (_, lo, hi) in enumerate(ranges):
        scores_all, match_all, ignore_all, npig = [], [], [], 0
        for img in image_ids:
            g_img = gts.get(img, [])
            d_img = sorted(dts.get(img, []), key=lambda r: -r["score"])[:max_dets]
            g_ign = np.array([bool(g.get("iscrowd", 0)) or g.get("num_keypoints", 1) == 0 or
                              g["area"] < lo or g["area"] > hi for g in g_img], dtype=bool)
            order = np.argsort(g_ign, kind="mergesort")               # unignored ground truths first
            g_img = [g_img[i] for i in order]
            g_ign = g_ign[order]
            npig += int((~g_ign).sum())
            sim = np.zeros((len(d_img), len(g_img)))
            for i, d in enumerate(d_img):
                for j, g in enumerate(g_img):
                    sim[i, j] = oks(g["keypoints"], g["area"], d["keypoints"], g.get("bbox"))
            d_area = [float(d["bbox"][2] * d["bbox"][3]) if "area" not in d else d["area"] for d in d_img]
            dm = -np.ones((len(thrs), len(d_img)), dtype=int)
            dig = np.zeros((len(thrs), len(d_img)), dtype=bool)
            for t, thr in enumerate(thrs):
                gm = -np.ones(len(g_img), dtype=int)
                for i in range(len(d_img)):
                    best, m = min(thr, 1 - 1e-10), -1
                    for j in range(len(g_img)):
                        if gm[j] >= 0 and not g_img[j].get("iscrowd", 0):
                            continue
                        if m > -1 and not g_ign[m] and g_ign[j]:
                            break                                      # only ignored ones remain
                        if sim[i, j] < best:
                            continue
                        best, m = sim[i, j], j
                    if m == -1:
                        continue
                    dig[t, i] = g_ign[m]
                    dm[t, i] = m
                    gm[m] = i
                for i in range(len(d_img)):                            # unmatched, outside the area range
                    if dm[t, i] == -1 and (d_area[i] < lo or d_area[i] > hi):
                        dig[t, i] = True
            scores_all.append([d["score"] for d in d_img])
            match_all.append(dm)
            ignore_all.append(dig)
        if npig == 0:
            continue
        sc = np.concatenate([np.asarray(s, dtype=np.float64) for s in scores_all]) if scores_all else np.zeros(0)
        inds = np.argsort(-sc, kind="mergesort")
        dm = np.concatenate(match_all, axis=1)[:, inds] if match_all else np.zeros((len(thrs), 0), int)
        dig = np.concatenate(ignore_all, axis=1)[:, inds] if ignore_all else np.zeros((len(thrs), 0), bool)
        tps = np.logical_and(dm >= 0, ~dig)
        fps = np.logical_and(dm < 0, ~dig)
        tp_sum = np.cumsum(tps, axis=1).astype(np.float64)
        fp_sum = np.cumsum(fps, axis=1).astype(np.float64)
        for t in range(len(thrs)):
            tp, fp = tp_sum[t], fp_sum[t]
            nd = len(tp)
            rc = tp / npig
            pr = tp / (fp + tp + np.spacing(1))
            recall[t, a] = rc[-1] if nd else 0.0
            pr = pr.tolist()
            for i in range(nd - 1, 0, -1):                             # monotone envelope
                if pr[i] > pr[i - 1]:
                    pr[i - 1] = pr[i]
            q = np.zeros(len(recs))
            idx = np.searchsorted(rc, recs, side="left")
            for ri, pi in enumerate(idx):
                if pi < nd:
                    q[ri] = pr[pi]
            precision[t, :, a] = q

    def _ap(t_sel=None, a=0):
        p = precision[:, :, a] if t_sel is None else precision[t_sel:t_sel + 1, :, a]
        p = p[p > -1]
        return float(np.mean(p)) if p.size else -1.0

    def _ar(t_sel=None, a=0):
        r = recall[:, a] if t_sel is None else recall[t_sel:t_sel + 1, a]
        r = r[r > -1]
        return float(np.mean(r)) if r.size else -1.0

    vals = [_ap(), _ap(0), _ap(5), _ap(None, 1), _ap(None, 2), _ar(), _ar(0), _ar(5), _ar(None, 1),
            _ar(None, 2)]
    return OrderedDict(zip(STATS_NAMES, vals))
