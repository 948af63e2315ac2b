"""Seeded synthetic inputs for the decode path.

Per image i the heat-map background is U(0, 0.05) and the tag background is
N(0, 0.05); P ~ U{1..max_people} people are planted with a centre U(60, 260)^2
(scaled to the map size), a per-joint offset U(-40, 40)^2, a sigma=2 Gaussian blob
of amplitude U(0.3, 1.0) max-merged into the heat map, and tag 1.5*p + N(0, 0.05)
wherever the blob exceeds 0.01.  Continuous fp32 noise keeps peak values distinct,
so the top-k order is well defined.
"""
from __future__ import annotations

import torch


def synth_decode_batch(num_images: int, num_joints: int = 17, height: int = 320,
                       width: int = 320, tag_dims: int = 1, max_people: int = 30,
                       seed: int = 1234, device="cpu", tag_per_joint: bool = True,
                       first_index: int = 0):
    """-> det (N,J,H,W) f32, tag (N,Jt,H,W,T) f32 on ``device``.

    Image i uses ``torch.Generator().manual_seed(seed + first_index + i)`` for the
    planted people, so a shard of a larger batch reproduces the same images."""
    dev = torch.device(device)
    jt = num_joints if tag_per_joint else 1
    gdev = torch.Generator(device=dev)
    gdev.manual_seed(seed * 7919 + first_index)
    det = torch.rand((num_images, num_joints, height, width), generator=gdev,
                     device=dev, dtype=torch.float32) * 0.05
    tag = torch.randn((num_images, jt, height, width, tag_dims), generator=gdev,
                      device=dev, dtype=torch.float32) * 0.05
    ys = torch.arange(height, device=dev, dtype=torch.float32)[:, None]
    xs = torch.arange(width, device=dev, dtype=torch.float32)[None, :]
    sy, sx = height / 320.0, width / 320.0
    for i in range(num_images):
        g = torch.Generator()
        g.manual_seed(seed + first_index + i)
        npeople = int(torch.randint(1, max_people + 1, (1,), generator=g))
        centre = 60.0 + 200.0 * torch.rand((npeople, 2), generator=g)
        offs = -40.0 + 80.0 * torch.rand((npeople, num_joints, 2), generator=g)
        amp = 0.3 + 0.7 * torch.rand((npeople, num_joints), generator=g)
        cx = ((centre[:, None, 0] + offs[:, :, 0]) * sx).clamp(2, width - 3).to(dev)
        cy = ((centre[:, None, 1] + offs[:, :, 1]) * sy).clamp(2, height - 3).to(dev)
        amp = amp.to(dev)
        bg = tag[i].clone()
        for p in range(npeople):
            blob = amp[p, :, None, None] * torch.exp(
                -((xs[None] - cx[p, :, None, None]) ** 2 +
                  (ys[None] - cy[p, :, None, None]) ** 2) / (2.0 * 2.0 ** 2))
            det[i] = torch.maximum(det[i], blob)
            mask = blob > 0.01                                   # (J,H,W)
            if tag_per_joint:
                tag[i] = torch.where(mask[..., None], 1.5 * p + bg, tag[i])
            else:
                any_mask = mask.any(dim=0)
                tag[i, 0] = torch.where(any_mask[..., None], 1.5 * p + bg[0], tag[i, 0])
    return det.contiguous(), tag.contiguous()
