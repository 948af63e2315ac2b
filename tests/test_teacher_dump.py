"""Teacher-dump writer / reader against the reference's own lines:
``np.savez_compressed(...)`` of teacher_inference.py:86-90 and ``_get_teacher_data`` of
rtpe/dataloaders.py:140-165 (restated inline: both are five lines of numpy / torch)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rtpe_b200 import teacher_dump as TD


def reference_write(out_path, preds, refined):
    # teacher_inference.py:86-90, verbatim semantics
    np.savez_compressed(out_path, pred_heatmaps=preds[:17], embeddings=preds[17:],
                        heatmaps_refined=refined, heatmaps_order=TD.HEATMAPS_ORDER)


def reference_read(teacher_dir, img_id, out_hw=None):
    # rtpe/dataloaders.py:148-165
    npz = np.load(os.path.join(teacher_dir, img_id + ".jpg_w48_predictions.npz"))
    t_hms = torch.FloatTensor(npz["heatmaps_refined"])
    t_ae = torch.FloatTensor(npz["embeddings"])
    if out_hw is not None:
        t_hms = F.interpolate(t_hms.unsqueeze(0), out_hw, mode="bilinear", align_corners=True)[0]
        t_ae = F.interpolate(t_ae.unsqueeze(0), out_hw, mode="bilinear", align_corners=True)[0]
    return t_hms, t_ae


def test_writer_matches_reference_format(tmp_path):
    rng = np.random.default_rng(0)
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    os.makedirs(ref)
    names = ["/data/coco/%012d.jpg" % i for i in range(7)]
    maps = [(rng.standard_normal((34, 16, 24), dtype=np.float32),
             rng.standard_normal((17, 32, 48), dtype=np.float32)) for _ in names]
    with TD.TeacherDumpWriter(str(ours), workers=3, max_pending=2) as w:
        for n, (p, r) in zip(names, maps):
            w.submit(n, p, r)
    assert w.written == len(names)
    for n, (p, r) in zip(names, maps):
        reference_write(os.path.join(str(ref), os.path.basename(n)) + "_w48_predictions", p, r)
        a = np.load(TD.dump_path(str(ours), n) + ".npz")
        b = np.load(TD.dump_path(str(ref), n) + ".npz")
        assert sorted(a.files) == sorted(b.files) == ["embeddings", "heatmaps_order",
                                                       "heatmaps_refined", "pred_heatmaps"]
        for k in b.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape
            assert np.array_equal(a[k], b[k])


def test_writer_reports_errors(tmp_path):
    w = TD.TeacherDumpWriter(str(tmp_path), workers=1)
    with pytest.raises(ValueError):
        w.submit("x.jpg", np.zeros((33, 4, 4), np.float32), np.zeros((17, 8, 8), np.float32))
    w.close()
    assert TD.load_teacher_data(None, "x")[0].numel() == 0


def test_reader_without_resize(tmp_path):
    rng = np.random.default_rng(1)
    p, r = rng.standard_normal((34, 8, 8), dtype=np.float32), rng.standard_normal((17, 16, 16), dtype=np.float32)
    reference_write(str(tmp_path / "000000000009.jpg_w48_predictions"), p, r)
    hm, ae = TD.load_teacher_data(str(tmp_path), "000000000009")
    rh, ra = reference_read(str(tmp_path), "000000000009")
    assert torch.equal(hm, rh) and torch.equal(ae, ra)
    with pytest.raises(Exception):
        TD.load_teacher_data(str(tmp_path), "000000000009", out_hw=(4, 4), device="cpu")


@pytest.mark.gpu
def test_reader_resize_gpu(cuda_device, tmp_path):
    rng = np.random.default_rng(2)
    p, r = rng.standard_normal((34, 40, 56), dtype=np.float32), rng.standard_normal((17, 80, 112), dtype=np.float32)
    reference_write(str(tmp_path / "000000000025.jpg_w48_predictions"), p, r)
    hm, ae = TD.load_teacher_data(str(tmp_path), "000000000025", out_hw=(123, 200), device=cuda_device)
    rh, ra = reference_read(str(tmp_path), "000000000025", out_hw=(123, 200))
    assert hm.shape == rh.shape and ae.shape == ra.shape
    # fp32 bilinear on both sides; only the FMA contraction order differs
    assert (hm.cpu() - rh).abs().max() <= 1e-5 * rh.abs().max()
    assert (ae.cpu() - ra).abs().max() <= 1e-5 * ra.abs().max()


@pytest.mark.gpu
def test_dumper_batches_gpu(cuda_device, tmp_path):
    """files written from batched forwards == the network outputs of each image, in the
    reference's layout; two batches exercise the pinned double buffers."""
    import rtpe_b200
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None).to(cuda_device)
    model[1].freeze()
    x = torch.randn(6, 3, 64, 96, device=cuda_device)
    names = ["img_%d.jpg" % i for i in range(6)]
    with TD.TeacherDumpWriter(str(tmp_path), workers=2) as w:
        d = TD.TeacherDumper(model, w)
        for b in range(3):
            d.dump_batch(x[2 * b:2 * b + 2], names[2 * b:2 * b + 2])
        d.close()
    with torch.no_grad():
        outs = [model(x[2 * b:2 * b + 2]) for b in range(3)]      # same batching as the dumper
        y0 = torch.cat([o[0].float() for o in outs], 0)
        y1 = torch.cat([o[1].float() for o in outs], 0)
    for i, n in enumerate(names):
        f = np.load(TD.dump_path(str(tmp_path), n) + ".npz")
        assert f["pred_heatmaps"].shape == (17, 16, 24) and f["heatmaps_refined"].shape == (17, 32, 48)
        assert list(f["heatmaps_order"]) == TD.HEATMAPS_ORDER
        got = np.concatenate([f["pred_heatmaps"], f["embeddings"]], 0)
        assert np.array_equal(got, y0[i].float().cpu().numpy())
        assert np.array_equal(f["heatmaps_refined"], y1[i].float().cpu().numpy())
