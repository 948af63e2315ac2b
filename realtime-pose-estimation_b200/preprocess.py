"""Image pre-processing of the teacher-inference driver on the GPU.

Reference (``teacher_inference.py:70-79``): PIL image -> ``resize_align_multi_scale(np.array(img),
640, 1, 1)`` (``rtpe/third_party/transforms.py:155-192``: target size rounded up to multiples of
64, affine matrix, ``cv2.warpAffine``) -> ``ToTensor`` -> ``Normalize``.  Here the host computes
the sizes and the 2x3 matrix with the reference's formulas, and ONE kernel
(``brtpe_preprocess_warp_normalize``) warps the uint8 image with OpenCV's fixed-point arithmetic
and writes the normalised float32 CHW tensor straight into the network's input batch -- byte
identical to the reference's CPU path (tests/test_preprocess.py).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

MEAN = (0.485, 0.456, 0.406)   # teacher_inference.py:72-73
STD = (0.229, 0.224, 0.225)


def get_multi_scale_size(image_hw, input_size, current_scale, min_scale):
    """transforms.py:155-176 -> ((w_resized, h_resized), center, scale)."""
    h, w = int(image_hw[0]), int(image_hw[1])
    center = np.array([int(w / 2.0 + 0.5), int(h / 2.0 + 0.5)])
    min_input_size = int((min_scale * input_size + 63) // 64 * 64)
    if w < h:
        w_resized = int(min_input_size * current_scale / min_scale)
        h_resized = int(int((min_input_size / w * h + 63) // 64 * 64) * current_scale / min_scale)
        scale_w = w / 200.0
        scale_h = h_resized / w_resized * w / 200.0
    else:
        h_resized = int(min_input_size * current_scale / min_scale)
        w_resized = int(int((min_input_size / h * w + 63) // 64 * 64) * current_scale / min_scale)
        scale_h = h / 200.0
        scale_w = w_resized / h_resized * h / 200.0
    return (w_resized, h_resized), center, np.array([scale_w, scale_h])


def get_affine_transform(center, scale, output_size):
    """transforms.py:59-93 for rot = 0, shift = 0, inv = 0.  The reference solves
    ``cv2.getAffineTransform`` for three float32 point pairs that describe an isotropic scale
    s = dst_w / (200 * scale[0]) about ``center``; this is that solution in closed form."""
    src_w = np.float32(scale[0] * 200.0)
    dst_w, dst_h = output_size[0], output_size[1]
    cx, cy = np.float32(center[0]), np.float32(center[1])
    sy1 = np.float32(cy + np.float32(src_w * np.float32(-0.5)))
    dcx, dcy = np.float32(dst_w * 0.5), np.float32(dst_h * 0.5)
    dy1 = np.float32(dcy + np.float32(np.float32(dst_w) * np.float32(-0.5)))
    s = (float(dy1) - float(dcy)) / (float(sy1) - float(cy))
    return np.array([[s, 0.0, float(dcx) - s * float(cx)],
                     [0.0, s, float(dcy) - s * float(cy)]], dtype=np.float64)


def get_inverse_affine_transform(center, scale, output_size):
    """transforms.py:59-93 with inv = 1 (``cv2.getAffineTransform(dst, src)``): maps heat-map /
    network-input coordinates back to the original image."""
    m = get_affine_transform(center, scale, output_size)
    s = m[0, 0]
    return np.array([[1.0 / s, 0.0, -m[0, 2] / s], [0.0, 1.0 / s, -m[1, 2] / s]], dtype=np.float64)


def transform_preds(coords, center, scale, output_size):
    """transforms.py:50-56: columns 0:2 of ``coords`` (P, >=2) through the inverse transform, the
    other columns (score, tags) unchanged; returns a copy like the reference."""
    target = coords.copy()
    t = get_inverse_affine_transform(center, scale, output_size)
    if coords.shape[0]:
        xy1 = np.concatenate([coords[:, 0:2].astype(np.float64),
                              np.ones((coords.shape[0], 1))], 1)
        target[:, 0:2] = xy1 @ t.T
    return target


def get_final_preds(grouped_joints, center, scale, heatmap_size):
    """transforms.py:195-202: ``grouped_joints = [list of (J, 3+T) person arrays]`` (the first
    element of ``HeatmapParser.parse``'s result) -> one transformed array per person."""
    return [transform_preds(person, center, scale, heatmap_size) for person in grouped_joints[0]]


def _as_device_u8(image, device):
    if isinstance(image, np.ndarray):
        image = torch.from_numpy(np.ascontiguousarray(image))
    if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
        raise ValueError("image must be uint8 (H, W, 3), got %s %s" % (image.dtype, tuple(image.shape)))
    return image.to(device, non_blocking=True).contiguous()


def warp_normalize(image, trans, dsize, out=None, want_u8=False, mean=MEAN, std=STD, device="cuda"):
    """``Normalize(ToTensor(cv2.warpAffine(image, trans, dsize)))`` -> float32 (3, Ho, Wo) on the
    GPU (written into ``out`` when given, e.g. one slot of the network's input batch); with
    ``want_u8`` also returns the warped uint8 (Ho, Wo, 3) image."""
    lib = L.load()
    dev = torch.device(device) if out is None else out.device
    img = _as_device_u8(image, dev)
    wo, ho = int(dsize[0]), int(dsize[1])
    if out is None:
        out = torch.empty((3, ho, wo), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (3, ho, wo) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 (3, %d, %d) tensor" % (ho, wo))
    u8 = torch.empty((ho, wo, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    m = (C.c_double * 6)(*np.asarray(trans, np.float64).reshape(6))
    mean_c = (C.c_float * 3)(*mean)
    std_c = (C.c_float * 3)(*std)
    with torch.cuda.device(dev):
        L.check(lib.brtpe_preprocess_warp_normalize(
            L.ptr(img), img.shape[0], img.shape[1], img.stride(0), m, ho, wo, mean_c, std_c,
            L.ptr(out), L.ptr(u8) if u8 is not None else None, L.stream_ptr(dev)),
            "brtpe_preprocess_warp_normalize")
    return (out, u8) if want_u8 else out


def resize_align_multi_scale(image, input_size, current_scale, min_scale, out=None, device="cuda"):
    """transforms.py:179-192 fused with ToTensor + Normalize:
    -> (float32 (3, h_resized, w_resized) CUDA tensor, center, scale)."""
    size, center, scale = get_multi_scale_size(image.shape[:2], input_size, current_scale, min_scale)
    trans = get_affine_transform(center, scale, size)
    return warp_normalize(image, trans, size, out=out, device=device), center, scale


def preprocess_image(image, input_size=640, device="cuda"):
    """teacher_inference.py:75-79: uint8 RGB (H, W, 3) -> (1, 3, H', W') network input."""
    t, center, scale = resize_align_multi_scale(image, input_size, 1, 1, device=device)
    return t.unsqueeze(0), center, scale
