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


def _affine_points(center, scale, output_size):
    """The three float32 point pairs of transforms.py:59-90 (rot = 0, shift = 0): the centre, the point
    half a source width above it and their perpendicular third point (``get_3rd_point``)."""
    src_w = np.asarray(scale, np.float64)[0] * 200.0
    dst_w, dst_h = output_size[0], output_size[1]
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = np.asarray(center)
    src[1, :] = np.asarray(center) + np.array([0.0, src_w * -0.5])
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)
    for p in (src, dst):
        d = p[0, :] - p[1, :]
        p[2, :] = p[1, :] + np.array([-d[1], d[0]], dtype=np.float32)
    return src, dst


def _solve_affine(src, dst):
    """``cv2.getAffineTransform(src, dst)``: the 2x3 matrix through three point pairs.  With OpenCV
    present (it is wherever the reference runs) its own solver is called, so the matrix is the
    reference's bit for bit -- float32 points make it slightly anisotropic, which a closed-form
    isotropic scale does not reproduce; without it the same 6x6 system is solved in float64."""
    try:
        import cv2
        return cv2.getAffineTransform(np.float32(src), np.float32(dst))
    except ImportError:
        a = np.zeros((6, 6))
        b = np.zeros(6)
        for i in range(3):
            a[2 * i, 0:2], a[2 * i, 2] = src[i], 1.0
            a[2 * i + 1, 3:5], a[2 * i + 1, 5] = src[i], 1.0
            b[2 * i], b[2 * i + 1] = dst[i]
        return np.linalg.solve(a, b).reshape(2, 3)


def get_affine_transform(center, scale, output_size):
    """transforms.py:59-93 for rot = 0, shift = 0, inv = 0."""
    src, dst = _affine_points(center, scale, output_size)
    return np.asarray(_solve_affine(src, dst), np.float64)


def get_inverse_affine_transform(center, scale, output_size):
    """transforms.py:59-93 with inv = 1 (``cv2.getAffineTransform(dst, src)``): maps heat-map /
    network-input coordinates back to the original image."""
    src, dst = _affine_points(center, scale, output_size)
    return np.asarray(_solve_affine(dst, src), np.float64)


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
