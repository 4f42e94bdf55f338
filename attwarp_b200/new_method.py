"""Drop-in mirror of the reference's NumPy warp API, running on the B200.

Mirrors ``Attention Guided Warping/new_method.py`` of dwipddalal/AttWarp:

* ``warp_image_by_attention(image, att_map, new_width, new_height)``  (new_method.py:198-283)
* ``set_transform_function(name, exp_scale, exp_divisor, apply_inverse)`` (new_method.py:378-403)
* ``save_warped_image(...) -> bool``                                  (new_method.py:405-506)
* ``resize_image_to_match_attmap``                                    (new_method.py:355-376)

Same signatures, argument meaning and error behaviour.  The arithmetic (float64 marginals /
CDF / inverse-CDF, cv2-compatible bilinear resample) runs in libattwarp_sm100.so through
``attwarp_warp_image_host``; there is no CPU fallback.  The module-level transform state of the
reference is kept for compatibility (``set_transform_function``), but every function also
accepts the transform explicitly, which is re-entrant.
"""

from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib

# module-level state mirroring new_method.py:159-195 (default transform is sqrt, :191)
ATTENTION_TRANSFORM = "sqrt"
EXP_SCALE = 1.0
EXP_DIVISOR = 1.0
APPLY_INVERSE_TO_MARGINALS = False
EPSILON = 1e-9
BASE_ATTENTION = 1e-9

_NP_DTYPE_IDS = {np.dtype(np.uint8): _lib.U8, np.dtype(np.float32): _lib.F32,
                 np.dtype(np.float64): _lib.F64}


def set_transform_function(transform_name, exp_scale=1.0, exp_divisor=1.0, apply_inverse=False):
    """Select the process-wide transform; unknown names fall back to identity and return
    ``"identity"`` (new_method.py:378-403)."""
    global ATTENTION_TRANSFORM, EXP_SCALE, EXP_DIVISOR, APPLY_INVERSE_TO_MARGINALS
    EXP_SCALE = exp_scale
    EXP_DIVISOR = exp_divisor
    APPLY_INVERSE_TO_MARGINALS = apply_inverse
    if transform_name in _lib.TRANSFORM_IDS:
        ATTENTION_TRANSFORM = transform_name
        return transform_name
    print(f"Unknown transform: {transform_name}. Using identity transform.")
    ATTENTION_TRANSFORM = "identity"
    return "identity"


def _coerce_att(att_map):
    att = np.asarray(att_map)
    if att.dtype not in _NP_DTYPE_IDS:
        # the reference does att_map.astype(np.float64) (new_method.py:207); float16 -> float32
        # and every integer/bool type -> float64 are exact, so the result is unchanged
        att = att.astype(np.float32 if att.dtype == np.float16 else np.float64)
    return np.ascontiguousarray(att)


def warp_image_by_attention(image, att_map, new_width, new_height, *, transform=None,
                            exp_scale=None, exp_divisor=None, apply_inverse=None):
    """Warp ``image`` ([h,w] or [h,w,c], uint8 or float32) by ``att_map`` ([h,w]).

    Positional signature identical to new_method.py:198.  With the keyword arguments left at
    ``None`` the module-level state set by ``set_transform_function`` is used, like the reference.
    """
    lib = _lib.load()
    image = np.asarray(image)
    if image.dtype not in (np.uint8, np.float32):
        raise TypeError(f"warp_image_by_attention: image dtype {image.dtype} not supported "
                        "(uint8 and float32 are)")
    h, w = image.shape[:2]
    c = 1 if image.ndim == 2 else image.shape[2]
    att = _coerce_att(att_map)
    if att.ndim != 2 or att.shape != (h, w):
        raise ValueError(f"att_map shape {att.shape} must equal the image's (h, w) = {(h, w)}")
    new_width, new_height = int(new_width), int(new_height)
    tp = _lib.make_transform(ATTENTION_TRANSFORM if transform is None else transform,
                             EXP_SCALE if exp_scale is None else exp_scale,
                             EXP_DIVISOR if exp_divisor is None else exp_divisor,
                             APPLY_INVERSE_TO_MARGINALS if apply_inverse is None else apply_inverse)
    img_c = np.ascontiguousarray(image)
    # cv2.remap returns [h,w] for a single-channel [h,w,1] input
    out_shape = (new_height, new_width) if c == 1 else (new_height, new_width, c)
    out = np.empty(out_shape, dtype=image.dtype)
    fallback = C.c_int(0)
    _lib.check(lib.attwarp_warp_image_host(
        img_c.ctypes.data_as(C.c_void_p), _NP_DTYPE_IDS[img_c.dtype], c, h, w,
        att.ctypes.data_as(C.c_void_p), _NP_DTYPE_IDS[att.dtype], new_width, new_height,
        C.byref(tp), out.ctypes.data_as(C.c_void_p), C.byref(fallback)))
    if fallback.value:
        print("Warning: Total attention is near zero.", file=sys.stderr)   # new_method.py:232
    return out


def resize_image_to_match_attmap(image, att_map):
    """new_method.py:355-376 (host-side cv2.resize; not on the hot path)."""
    if image is None or att_map is None:
        return None
    import cv2
    target_h, target_w = att_map.shape[:2]
    if image.shape[:2] == (target_h, target_w):
        return image.copy()
    try:
        resized = cv2.resize(image, (target_w, target_h), interpolation=cv2.INTER_LINEAR)
        if resized.shape[:2] != (target_h, target_w):
            raise RuntimeError("Resize resulted in unexpected shape.")
        return resized
    except Exception as e:  # noqa: BLE001 - mirrors the reference's catch-all
        print(f"Error resizing image: {e}")
        return None


def save_warped_image(image_path, att_map, original_image_save_path, masked_overlay_save_path,
                      output_path, vis_path=None, width=500, height=500, transform="identity",
                      exp_scale=1.0, exp_divisor=1.0, apply_inverse=False, attention_alpha=0.5):
    """Load / coerce, resize the image to the attention-map size, warp on the GPU, write PNGs.
    Never raises: prints and returns ``False`` on any failure (new_method.py:405-506).
    File I/O and the overlay stay on the host with OpenCV exactly like the reference."""
    try:
        import cv2
        from PIL import Image
        if isinstance(image_path, str):
            image = cv2.imread(image_path)
            if image is None:
                raise ValueError(f"Could not read image: {image_path}")
        else:
            image = cv2.cvtColor(np.array(image_path), cv2.COLOR_RGB2BGR)
        in_h, in_w = image.shape[:2]
        if original_image_save_path:
            cv2.imwrite(original_image_save_path, image.copy())

        if isinstance(att_map, Image.Image):
            att_map = np.array(att_map)
        elif isinstance(att_map, list):
            if len(att_map) > 0:
                first = att_map[0]
                att_map = first if isinstance(first, np.ndarray) else np.array(first)
            else:
                att_map = np.ones((height, width), dtype=np.float32) * 128
        if att_map.ndim == 3:
            att_map = np.mean(att_map, axis=2)
        elif att_map.ndim != 2:
            raise ValueError(f"Attention map must be 2D, got shape {att_map.shape}")

        if masked_overlay_save_path:
            base = image.copy()
            if base.ndim == 2:
                base = cv2.cvtColor(base, cv2.COLOR_GRAY2BGR)
            att_r = cv2.resize(att_map.copy(), (in_w, in_h), interpolation=cv2.INTER_LINEAR)
            lo, hi = np.min(att_r), np.max(att_r)
            att_n = (att_r - lo) / (hi - lo) if hi > lo + EPSILON else np.zeros_like(att_r)
            heat = cv2.applyColorMap((att_n * 255).astype(np.uint8), cv2.COLORMAP_JET)
            cv2.imwrite(masked_overlay_save_path,
                        cv2.addWeighted(heat, attention_alpha, base, 1 - attention_alpha, 0))

        image_for_warping = resize_image_to_match_attmap(image, att_map)
        if image_for_warping is None:
            raise ValueError("Failed to resize image to match attention map dimensions for warping")
        transform_name = set_transform_function(transform, exp_scale, exp_divisor, apply_inverse)
        warped = warp_image_by_attention(image_for_warping, att_map, width, height)
        cv2.imwrite(output_path, warped)
        if vis_path:
            _write_visualization(image_for_warping, att_map, warped, vis_path, transform_name,
                                 attention_alpha)
        return True
    except Exception as e:  # noqa: BLE001 - the reference swallows everything (new_method.py:504-506)
        print(f"Error during processing: {e}")
        return False


def _write_visualization(image, att_map, warped, path, transform_name, attention_alpha):
    """Three-panel strip (input | attention overlay | warped), cf. new_method.py:285-353."""
    import cv2
    lo, hi = np.min(att_map), np.max(att_map)
    att_n = (att_map - lo) / (hi - lo) if hi > lo + EPSILON else np.zeros_like(att_map, dtype=np.float64)
    heat = cv2.applyColorMap((att_n * 255).astype(np.uint8), cv2.COLORMAP_JET)
    h = max(image.shape[0], warped.shape[0])

    def fit(im):
        if im.ndim == 2:
            im = cv2.cvtColor(im, cv2.COLOR_GRAY2BGR)
        s = h / im.shape[0]
        return cv2.resize(im, (max(int(round(im.shape[1] * s)), 1), h))

    base = image if image.ndim == 3 else cv2.cvtColor(image, cv2.COLOR_GRAY2BGR)
    overlay = cv2.addWeighted(heat, attention_alpha, base, 1 - attention_alpha, 0)
    strip = np.concatenate([fit(base), fit(overlay), fit(warped)], axis=1)
    cv2.putText(strip, f"transform: {transform_name}", (8, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.6,
                (255, 255, 255), 1, cv2.LINE_AA)
    cv2.imwrite(path, strip)
