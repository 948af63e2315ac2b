"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's image pre-processing
(SURVEY.md section 8f, rank 2); imported by tests/ only.

Reference chain (``teacher_inference.py:70-79``):

    resized, center, scale = resize_align_multi_scale(np.array(img), 640, 1, 1)
    t = Compose([ToTensor(), Normalize(mean, std)])(resized).unsqueeze(0)

* ``get_multi_scale_size`` / ``resize_align_multi_scale``: ``rtpe/third_party/transforms.py:155-192``
* ``get_affine_transform`` (rot = 0, shift = 0): ``transforms.py:59-93``, ``get_dir`` ``:107-114``,
  ``get_3rd_point`` ``:102-104``; the reference calls ``cv2.getAffineTransform`` on three float32
  point pairs -- for rot = 0 these describe an isotropic scale + translation, restated in closed
  form below (cv2's LU solve returns the same matrix up to ~1e-16 noise).
* ``cv2.warpAffine(image, trans, size)``: OpenCV is a third-party dependency that is not vendored
  in the reference; version in this image: 4.13.0.  Its 8-bit INTER_LINEAR / BORDER_CONSTANT path is
  fixed point (imgwarp.cpp ``WarpAffineInvoker`` + ``remapBilinear``): the matrix is inverted in
  double, source coordinates are rounded to 1/32 pixel with 10 fractional bits
  (``AB_BITS`` = 10, ``INTER_BITS`` = 5, ``round_delta`` = 16), the four weights are
  ``(32-fx)(32-fy)*32 ...`` (sum 2^15) and the result is ``(sum + 2^14) >> 15``.
* ``ToTensor`` + ``Normalize``: float32 ``(u8 / 255 - mean) / std`` (two IEEE divisions).

Pinned: ``oracle/make_golden_preprocess.py`` runs the UNMODIFIED reference functions + cv2 +
torchvision in this container and commits ``tests/golden/preprocess_*.npz``; ``warp_affine_u8``
reproduces ``cv2.warpAffine`` bit for bit on them and on the two bundled JPEGs at full size
(checked by the generating script), and ``tests/test_preprocess.py`` re-checks against cv2 live.
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def get_multi_scale_size(h, w, input_size, current_scale, min_scale):
    """transforms.py:155-176 -> ((w_resized, h_resized), center (2,) int, scale (2,) float64)."""
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
    """transforms.py:59-93 with rot = 0, shift = 0, inv = 0: the three point pairs are
    c -> (dw/2, dh/2), c + (0, -sw/2) -> (dw/2, dh/2 - dw/2) and their perpendicular third points,
    i.e. x' = s (x - cx) + dw/2, y' = s (y - cy) + dh/2 with s = dw / sw, sw = scale[0] * 200.
    The reference rounds the points to float32 before solving (``np.float32(src)``)."""
    src_w = np.float32(scale[0] * 200.0)
    dst_w, dst_h = output_size[0], output_size[1]
    cx, cy = np.float32(center[0]), np.float32(center[1])
    # float32 points, as cv2.getAffineTransform receives them
    sy1 = np.float32(cy + np.float32(src_w * np.float32(-0.5)))
    dcx, dcy = np.float32(dst_w * 0.5), np.float32(dst_h * 0.5)
    dy1 = np.float32(dcy + np.float32(np.float32(dst_w) * np.float32(-0.5)))
    s = (float(dy1) - float(dcy)) / (float(sy1) - float(cy))
    return np.array([[s, 0.0, float(dcx) - s * float(cx)],
                     [0.0, s, float(dcy) - s * float(cy)]], dtype=np.float64)


def invert_affine(m):
    """imgwarp.cpp cv::warpAffine: inversion of the 2x3 matrix in double."""
    m = np.asarray(m, np.float64).copy()
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    m[0, 0] = a11
    m[0, 1] *= -d
    m[1, 0] *= -d
    m[1, 1] = a22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2] = b1
    m[1, 2] = b2
    return m


def warp_affine_u8(img, trans, dsize):
    """cv2.warpAffine(img, trans, dsize) for uint8 HxWxC (INTER_LINEAR, BORDER_CONSTANT 0)."""
    wo, ho = int(dsize[0]), int(dsize[1])
    m = invert_affine(trans)
    x = np.arange(wo, dtype=np.float64)
    y = np.arange(ho, dtype=np.float64)
    adelta = np.rint(m[0, 0] * x * 1024).astype(np.int64)
    bdelta = np.rint(m[1, 0] * x * 1024).astype(np.int64)
    x0 = np.rint((m[0, 1] * y + m[0, 2]) * 1024).astype(np.int64) + 16
    y0 = np.rint((m[1, 1] * y + m[1, 2]) * 1024).astype(np.int64) + 16
    xx = (x0[:, None] + adelta[None, :]) >> 5
    yy = (y0[:, None] + bdelta[None, :]) >> 5
    sx, sy, fx, fy = xx >> 5, yy >> 5, xx & 31, yy & 31
    wts = [(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32]
    h, w = img.shape[:2]

    def tap(ty, tx):
        ok = (ty >= 0) & (ty < h) & (tx >= 0) & (tx < w)
        v = img[np.clip(ty, 0, h - 1), np.clip(tx, 0, w - 1)].astype(np.int64)
        return v * ok[..., None]

    acc = (tap(sy, sx) * wts[0][..., None] + tap(sy, sx + 1) * wts[1][..., None] +
           tap(sy + 1, sx) * wts[2][..., None] + tap(sy + 1, sx + 1) * wts[3][..., None])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def to_tensor_normalize(img_u8, mean=MEAN, std=STD):
    """torchvision ToTensor + Normalize: HxWx3 uint8 -> (3,H,W) float32."""
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    mean = np.asarray(mean, np.float32)[:, None, None]
    std = np.asarray(std, np.float32)[:, None, None]
    return ((x - mean) / std).astype(np.float32)


def resize_align_multi_scale(img, input_size, current_scale, min_scale):
    """transforms.py:179-192 -> (image_resized uint8, center, scale)."""
    size, center, scale = get_multi_scale_size(img.shape[0], img.shape[1], input_size,
                                               current_scale, min_scale)
    trans = get_affine_transform(center, scale, size)
    return warp_affine_u8(img, trans, size), center, scale
