"""``eval_student`` -- drop-in for rtpe/engine.py:21-75 (the evaluation loop of the distillation
scripts: model forward -> ``HeatmapParser.parse`` -> ``dataset.evaluate``).

Same signature, same per-image results and the same ``dataset.evaluate(all_preds, all_scores, ".",
False, False)`` call as the reference.  Differences, all on the device side:

* the prediction never travels to the host (the reference does ``pred.cpu()`` and hands numpy-backed
  CPU tensors to the parser, engine.py:42-49): heat-maps and embeddings stay CUDA tensors and the
  parser's kernels decode them in place;
* ``batched=True`` (default) decodes all images of a loader batch with ONE ``parse_batch`` call
  instead of the reference's implicit batch-of-one (``parse`` keeps only image 0 when ``refine=True``,
  group.py:278-287, which is why the reference's loaders use ``batch_size=1``); results are
  appended image by image, so the evaluation input is identical for ``batch_size=1`` loaders.

Plotting / JPEG dumps (``plot_every`` / ``save_every``: matplotlib + rtpe/third_party/vis.py) are
outside the hot path and raise ``NotImplementedError`` when requested.
"""
from __future__ import annotations

import torch


def _student_prediction(model, img, out_hw):
    """The reference calls ``model(img, out_hw)`` and expects one tensor (engine.py:41); the
    drop-in students return a tensor (``RefinerStudent``), a list of stage outputs
    (``CamStudent`` / ``MultistageStudent``: last stage) or ``(att, det)`` (attention students)."""
    pred = model(img, out_hw)
    if isinstance(pred, tuple):
        pred = pred[1]
    elif isinstance(pred, list):
        pred = pred[-1]
    return pred


def eval_student(model, hm_parser, val_dataloader, device, plot_every=None, save_every=None,
                 save_dir="/tmp", batched=True, num_heatmaps=17, verbose=True):
    """rtpe/engine.py:21-75.  ``val_dataloader`` yields ``(img_id, img, mask, hms, _, _)``;
    ``val_dataloader.dataset.evaluate`` computes the COCO metrics.  -> ``eval_dict``."""
    if plot_every is not None or save_every is not None:
        raise NotImplementedError("plotting / image dumps are outside the B200 hot path "
                                  "(rtpe/engine.py:52-67 uses matplotlib and third_party/vis.py)")
    model.eval()
    all_preds = []
    all_scores = []
    with torch.no_grad():
        for batch_i, batch in enumerate(val_dataloader):
            img = batch[1]
            if verbose:
                print("eval:", batch_i)
            out_hw = tuple(img.shape[2:])
            img = img.to(device)
            pred = _student_prediction(model, img, out_hw).detach()
            pred_hms = pred[:, :num_heatmaps]
            pred_ae = pred[:, num_heatmaps:].unsqueeze(-1)    # (N, AE_DIM, h, w, 1)
            if batched:
                for people, scores in hm_parser.parse_batch(pred_hms, pred_ae, adjust=True, refine=True):
                    all_preds.append([x for x in people if x.size > 0])
                    all_scores.append(scores)
            else:
                grouped, scores = hm_parser.parse(pred_hms, pred_ae, adjust=True, refine=True)
                all_preds.append([x for x in grouped[0] if x.size > 0])
                all_scores.append(scores)
    eval_dict, mAP = val_dataloader.dataset.evaluate(all_preds, all_scores, ".", False, False)
    if verbose:
        print("\n".join([k + "=" + str(v) for k, v in eval_dict.items()]))
    return eval_dict
