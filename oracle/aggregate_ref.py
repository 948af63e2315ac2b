"""ORACLE (test infrastructure only) -- heat-map / tag aggregation between the
network and the parser.

1. ``aggregate_intree_ref``: the in-tree simplified mode, validate_hhrnet.py:92-101
   (bilinear, align_corners=True, straight to the original image size, no flip,
   no stage average).
2. ``multi_stage_outputs_ref`` / ``aggregate_results_ref`` / ``aggregate_flip_multiscale_ref``:
   the flip-test + multi-scale mode.  Its only in-tree callers are
   legacy/valid_ae_avg.py:159-205 and legacy/valid_ae1dim.py:159-205, which import
   ``core.inference.get_multi_stage_outputs / aggregate_results`` from the upstream
   HigherHRNet repository -- NOT vendored under /root/reference, no pinned version.
   Restated from the published upstream algorithm with the
   configuration the reference keeps at legacy/distillation.py:85-92 (FLIP_TEST,
   PROJECT2IMAGE, WITH_HEATMAPS (True, True), WITH_AE (True, False), TAG_PER_JOINT).
   PARITY UNPINNED for this mode: no reference test or fixture exists for it.

Only tests/, smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# COCO left/right swap implied by the joint order at teacher_inference.py:38-40
FLIP_INDEX = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]


def aggregate_intree_ref(y0: torch.Tensor, y1: torch.Tensor, out_hw, num_joints=17):
    """-> det (N,J,h,w), tag (N,A,h,w,1)  (validate_hhrnet.py:94-101)."""
    det = F.interpolate(y1, out_hw, mode="bilinear", align_corners=True)
    tag = F.interpolate(y0[:, num_joints:], out_hw, mode="bilinear", align_corners=True)
    return det, tag.unsqueeze(-1)


def multi_stage_outputs_ref(outputs, outputs_flip=None, size_projected=None,
                            num_joints=17, tag_per_joint=True, flip_index=FLIP_INDEX):
    """upstream get_multi_stage_outputs with the model calls factored out:
    ``outputs`` = model(image), ``outputs_flip`` = model(flip(image, [3])) or None.
    -> heatmaps (list of 1 or 2 tensors), tags (list of 1 or 2 tensors)."""
    def one(outs, flipped):
        y0, y1 = outs
        up = F.interpolate(y0, size=(y1.shape[2], y1.shape[3]), mode="bilinear",
                           align_corners=False)
        hm = (up[:, :num_joints] + y1[:, :num_joints]) / 2.0
        tg = up[:, num_joints:]
        if flipped:
            hm = torch.flip(hm, [3])[:, flip_index]
            tg = torch.flip(tg, [3])
            if tag_per_joint:
                tg = tg[:, flip_index]
        return hm, tg

    hms, tgs = [], []
    h, t = one(outputs, False)
    hms.append(h)
    tgs.append(t)
    if outputs_flip is not None:
        h, t = one(outputs_flip, True)
        hms.append(h)
        tgs.append(t)
    if size_projected is not None:
        hw = (size_projected[1], size_projected[0])
        hms = [F.interpolate(v, size=hw, mode="bilinear", align_corners=False) for v in hms]
        tgs = [F.interpolate(v, size=hw, mode="bilinear", align_corners=False) for v in tgs]
    return hms, tgs


def aggregate_results_ref(scale, final_heatmaps, tags_list, heatmaps, tags, num_scales,
                          project2image=True):
    """upstream aggregate_results."""
    if scale == 1 or num_scales == 1:
        if final_heatmaps is not None and not project2image:
            tags = [F.interpolate(t, size=(final_heatmaps.size(2), final_heatmaps.size(3)),
                                  mode="bilinear", align_corners=False) for t in tags]
        for t in tags:
            tags_list.append(t.unsqueeze(4))
    avg = (heatmaps[0] + heatmaps[1]) / 2.0 if len(heatmaps) > 1 else heatmaps[0]
    if final_heatmaps is None:
        final_heatmaps = avg
    elif project2image:
        final_heatmaps = final_heatmaps + avg
    else:
        final_heatmaps = final_heatmaps + F.interpolate(
            avg, size=(final_heatmaps.size(2), final_heatmaps.size(3)),
            mode="bilinear", align_corners=False)
    return final_heatmaps, tags_list


def aggregate_flip_multiscale_ref(per_scale_outputs, base_size, num_joints=17,
                                  tag_per_joint=True):
    """``per_scale_outputs``: list of (scale, outputs, outputs_flip_or_None), visited in
    the given order (callers pass scales descending, legacy/valid_ae_avg.py:166).
    ``base_size`` = (W, H).  -> det (N,J,H,W), tag (N,A,H,W,T)."""
    final, tags_list = None, []
    ns = len(per_scale_outputs)
    for scale, outs, outs_flip in per_scale_outputs:
        hms, tgs = multi_stage_outputs_ref(outs, outs_flip, base_size, num_joints, tag_per_joint)
        final, tags_list = aggregate_results_ref(scale, final, tags_list, hms, tgs, ns)
    final = final / float(ns)
    return final, torch.cat(tags_list, dim=4)
