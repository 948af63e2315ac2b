"""COCO keypoint result records + OKS evaluation.  The record builder is
pinned against the reference's OWN method bodies (rtpe/third_party/COCODataset.py:135-275), pulled
out of the source file with ``ast`` because the module itself needs pycocotools / json_tricks /
cv2 at import time (this container only; skipped where /root/reference is absent).  The OKS
evaluation restates the published COCO protocol (pycocotools is not installed: parity unpinned) and
is checked on hand-computed cases."""
import ast
import json
import logging
import os
from collections import OrderedDict, defaultdict

import numpy as np
import pytest

from rtpe_b200 import coco_results as CR
from oracle.ref_loader import REF_ROOT, reference_available


def _people(rng, n_img=3, t=1):
    preds, scores, ids = [], [], []
    for i in range(n_img):
        k = int(rng.integers(0, 4)) if i != 1 else 0                       # image 1: no detection
        people = [np.concatenate([rng.uniform(0, 400, (17, 2)), rng.uniform(0, 1, (17, 1)),
                                  rng.normal(0, 1, (17, t))], 1).astype(np.float32) for _ in range(k)]
        preds.append(people)
        scores.append([np.float32(rng.uniform(0, 1)) for _ in range(k)])
        ids.append(1000 + 7 * i)
    return preds, scores, ids


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
def test_result_records_match_reference_methods(tmp_path):
    src = open(os.path.join(REF_ROOT, "rtpe", "third_party", "COCODataset.py")).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "CocoDataset")
    want = {"processKeypoints", "evaluate", "_write_coco_keypoint_results",
            "_coco_keypoint_results_one_category_kernel"}
    body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in body} == want
    mod = ast.Module(body=[ast.ClassDef(name="Ref", bases=[], keywords=[], body=body, decorator_list=[])],
                     type_ignores=[])
    ast.fix_missing_locations(mod)

    class NP:                                                              # numpy 2.x has no np.float
        float = float

        def __getattr__(self, name):
            return getattr(np, name)

    class PlainJson:                                                       # json_tricks stand-in
        @staticmethod
        def dump(obj, f, **kw):
            json.dump(obj, f, default=lambda o: o.item(), **kw)
        load = staticmethod(json.load)

    ns = {"np": NP(), "json": PlainJson, "os": os, "defaultdict": defaultdict, "OrderedDict": OrderedDict,
          "logger": logging.getLogger("ref")}
    exec(compile(mod, "COCODataset.py", "exec"), ns)
    rng = np.random.default_rng(5)
    preds, scores, ids = _people(rng)

    class FakeCoco:
        def loadImgs(self, img_id):
            return [{"file_name": "%012d.jpg" % img_id}]

    ref = ns["Ref"]()
    ref.dataset, ref.ids, ref.coco = "test2017", ids, FakeCoco()           # 'test': no COCOeval call
    ref.classes = ["__background__", "person"]
    ref._class_to_coco_ind = {"person": 1}
    out = ref.evaluate(preds, scores, str(tmp_path), False, False)
    assert out == ({"Null": 0}, 0)
    theirs = json.load(open(tmp_path / "results" / "keypoints_test2017_results.json"))
    mine = CR.build_keypoint_results(preds, scores, ids)
    CR.write_keypoint_results(mine, tmp_path / "mine.json")
    assert json.load(open(tmp_path / "mine.json")) == theirs
    assert open(tmp_path / "mine.json").read() == \
        open(tmp_path / "results" / "keypoints_test2017_results.json").read()
    assert len(mine) == sum(len(p) for p in preds) and all(len(r["keypoints"]) == 51 for r in mine)


def _gt(image_id, kp, area=5000.0, num=None, crowd=0):
    kp = np.asarray(kp, dtype=np.float64)
    x0, y0, x1, y1 = kp[:, 0].min(), kp[:, 1].min(), kp[:, 0].max(), kp[:, 1].max()
    return {"image_id": image_id, "keypoints": kp.reshape(-1).tolist(), "area": area,
            "num_keypoints": int((kp[:, 2] > 0).sum()) if num is None else num,
            "bbox": [x0, y0, x1 - x0, y1 - y0], "iscrowd": crowd}


def _det(image_id, kp, score):
    kp = np.asarray(kp, dtype=np.float64)
    return CR.build_keypoint_results([[np.concatenate([kp, np.zeros((17, 1))], 1)]], [[score]], [image_id])[0]


def test_oks_formula():
    rng = np.random.default_rng(1)
    kp = np.concatenate([rng.uniform(50, 250, (17, 2)), np.full((17, 1), 2.0)], 1)
    assert CR.oks(kp, 5000.0, kp) == pytest.approx(1.0)
    moved = kp.copy()
    moved[:, 0] += 10.0                                                   # every joint 10 px off
    want = np.mean(np.exp(-(10.0 ** 2) / ((2 * CR.COCO_SIGMAS) ** 2) / (5000.0 + np.spacing(1)) / 2))
    assert CR.oks(kp, 5000.0, moved) == pytest.approx(want, rel=1e-12)
    half = kp.copy()
    half[8:, 2] = 0                                                       # unlabelled joints do not count
    far = moved.copy()
    far[8:, :2] += 1e4
    assert CR.oks(half, 5000.0, far) == pytest.approx(
        np.mean(np.exp(-(10.0 ** 2) / ((2 * CR.COCO_SIGMAS[:8]) ** 2) / 5000.0 / 2)), rel=1e-9)
    none = kp.copy()
    none[:, 2] = 0                                                        # no label at all: doubled-box rule
    inside = np.concatenate([np.full((17, 2), 150.0), np.ones((17, 1))], 1)
    assert CR.oks(none, 5000.0, inside, gt_bbox=[100, 100, 100, 100]) == pytest.approx(1.0)


def test_evaluate_keypoints_hand_computed_cases():
    rng = np.random.default_rng(2)
    kp = [np.concatenate([rng.uniform(50, 250, (17, 2)), np.full((17, 1), 2.0)], 1) for _ in range(3)]
    gts = [_gt(1, kp[0]), _gt(2, kp[1])]
    perfect = CR.evaluate_keypoints(gts, [_det(1, kp[0], 0.9), _det(2, kp[1], 0.8)])
    assert perfect["AP"] == pytest.approx(1.0) and perfect["AR"] == pytest.approx(1.0)
    assert perfect["AP (M)"] == pytest.approx(1.0) and perfect["AP (L)"] == -1.0      # area 5000: medium
    # a higher-scored false positive in front of the two true positives: precision 1/2, 2/3 -> envelope 2/3
    fp = kp[2] + np.array([3000.0, 0, 0])
    worse = CR.evaluate_keypoints(gts, [_det(1, fp, 0.95), _det(1, kp[0], 0.9), _det(2, kp[1], 0.8)])
    assert worse["AP"] == pytest.approx(2.0 / 3.0) and worse["AR"] == pytest.approx(1.0)
    # one of two people found: precision 1 up to recall 0.5 (51 of the 101 recall points)
    half = CR.evaluate_keypoints(gts, [_det(1, kp[0], 0.9)])
    assert half["AP"] == pytest.approx(51.0 / 101.0) and half["AR"] == pytest.approx(0.5)
    # a detection 2 px off passes the loose thresholds only
    off = kp[0] + np.array([6.0, 0, 0])
    o = CR.oks(kp[0], 5000.0, off)
    loose = CR.evaluate_keypoints([gts[0]], [_det(1, off, 0.9)])
    passed = int((np.linspace(0.5, 0.95, 10) <= o + 1e-12).sum())
    assert 0 < passed < 10 and loose["AP"] == pytest.approx(passed / 10.0)
    assert loose["Ap .5"] == pytest.approx(1.0)
    # crowd / unlabelled ground truths are ignored, a detection matched to them is no false positive
    crowd = CR.evaluate_keypoints([gts[0], _gt(1, kp[2], crowd=1)], [_det(1, kp[0], 0.9), _det(1, kp[2], 0.5)])
    assert crowd["AP"] == pytest.approx(1.0)
    assert list(perfect.keys()) == CR.STATS_NAMES
