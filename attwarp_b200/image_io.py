"""On-disk formats either side of the warp: the reference drivers read JPEG files
(``Image.open(path).convert('RGB')``, AGW/main.py:152, main_batched.py:98), hand PIL images to ``save_warped_image``,
which flips them to BGR (new_method.py:421-422) and writes PNG files with ``cv2.imwrite`` (:491).  With the warp at
~0.1 ms per batch those two steps ARE the driver's run time, so this module moves the decode onto the GPU and
batches the rest:

* ``decode_jpeg_batch(sources)``     JPEG bytes / paths -> uint8 HWC **BGR** device tensors, the layout stage 5
  reads (nvJPEG through ``torchvision.io.decode_jpeg(device=...)`` -- library code, like calling cuBLAS).
  NOT bit-identical to libjpeg: IDCT and chroma up-sampling differ between decoders.  Measured against Pillow (the
  reader the reference drivers use) on the B200 box, synthetic photo-like images with a saturated colour block,
  quality 90 (tests/test_image_io.py holds these bounds): 4:4:4 files: max 4 LSB, mean 0.5 LSB; 4:2:0 files: mean
  1.9 LSB, 99 % of the bytes within ~10 LSB, but up to ~100 LSB on the one-pixel rim of a saturated colour edge
  (nvJPEG replicates chroma samples where libjpeg's "fancy" up-sampling interpolates them).  ``exact=True`` decodes
  on the host with Pillow instead -- bit-identical to ``Image.open(path).convert('RGB')`` -- for runs that must
  reproduce the reference's pixels.  Anything that is not a JPEG (PNG, ...) is decoded on the host, exactly.
* ``encode_png_batch(images, paths)`` device tensors -> PNG files: device -> pinned host copy, then
  ``cv2.imwrite`` in a thread pool -- the reference's own encoder, so the files decode to the same bytes.
  (A GPU PNG/deflate encoder is out of scope until its output is pinned; PNG is lossless, so only speed is at stake.)
* ``warp_files(...)``                the driver loop of main_batched.py:243-287 for a whole list of files: decode
  -> one ragged launch per stage -> encode.
"""

from __future__ import annotations

import concurrent.futures as cf
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops


def _read_bytes(src) -> bytes:
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    with open(os.fspath(src), "rb") as f:
        return f.read()


def _is_jpeg(buf: bytes) -> bool:
    return len(buf) > 3 and buf[0] == 0xFF and buf[1] == 0xD8


def decode_jpeg_batch(sources: Sequence, device=None, exact: bool = False) -> List[torch.Tensor]:
    """Paths or encoded bytes -> list of uint8 [H, W, 3] BGR tensors on ``device`` (cv2.imread's channel order, the
    one ``save_warped_image`` warps in).  JPEG files are decoded on the GPU in one batched nvJPEG call; grey JPEGs
    come out as three equal channels (like ``Image.convert('RGB')``).  ``exact=True``: host decode with Pillow,
    bit-identical to the reference's reader, then one copy to the device."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    bufs = [_read_bytes(s) for s in sources]
    out: List[Optional[torch.Tensor]] = [None] * len(bufs)
    if exact:
        import io

        from PIL import Image
        for i, b in enumerate(bufs):
            rgb = np.array(Image.open(io.BytesIO(b)).convert("RGB"))
            out[i] = torch.from_numpy(np.ascontiguousarray(rgb[..., ::-1])).to(dev, non_blocking=True)
        return out  # type: ignore[return-value]
    from torchvision.io import ImageReadMode, decode_jpeg
    jpeg_idx = [i for i, b in enumerate(bufs) if _is_jpeg(b)]
    if jpeg_idx:
        datas = [torch.frombuffer(bytearray(bufs[i]), dtype=torch.uint8) for i in jpeg_idx]
        decoded = decode_jpeg(datas, mode=ImageReadMode.RGB, device=dev)          # list of [3, H, W] RGB
        for i, t in zip(jpeg_idx, decoded):
            out[i] = t.flip(0).permute(1, 2, 0).contiguous()                      # -> [H, W, 3] BGR
    for i, b in enumerate(bufs):
        if out[i] is None:
            import cv2
            img = cv2.imdecode(np.frombuffer(b, dtype=np.uint8), cv2.IMREAD_COLOR)
            if img is None:
                raise ValueError(f"decode_jpeg_batch: source {i} is not a decodable image")
            out[i] = torch.from_numpy(img).to(dev)
    return out  # type: ignore[return-value]


def encode_png_batch(images: Sequence[torch.Tensor], paths: Sequence[str], workers: int = 8) -> List[bool]:
    """uint8 [H, W, 3] (BGR) or [H, W] device tensors -> PNG files with ``cv2.imwrite`` (the reference's encoder,
    new_method.py:491).  The device -> host copies are enqueued first (pinned staging), the encodes run in a
    thread pool (OpenCV releases the GIL).  Returns cv2.imwrite's result per file."""
    import cv2
    assert len(images) == len(paths)
    host = []
    for t in images:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        host.append(h)
    if images:
        torch.cuda.current_stream(images[0].device).synchronize()

    def write(k):
        return bool(cv2.imwrite(os.fspath(paths[k]), host[k].numpy()))

    if workers <= 1 or len(paths) <= 1:
        return [write(k) for k in range(len(paths))]
    with cf.ThreadPoolExecutor(max_workers=min(workers, len(paths))) as ex:
        return list(ex.map(write, range(len(paths))))


def warp_files(image_sources: Sequence, tok: torch.Tensor, output_paths: Sequence[str], out_sizes=None,
               transform: str = "identity", exp_scale: float = 1.0, exp_divisor: float = 1.0,
               apply_inverse: bool = False, device=None, workers: int = 8) -> List[bool]:
    """The per-image loop of the batched driver (main_batched.py:243-287: ``save_warped_image(image, mask, ...,
    width, height, "identity")`` per sample) for a list of image files and their token maps: GPU decode -> stages
    2-5 for the whole (mixed-resolution) list in one launch per stage -> PNG files.

    tok         [n, gh, gw] float32 token maps (the aggregated attention of each sample)
    out_sizes   per-image (height, width) of the warped files, default = the input sizes (the drivers pass the
                sample's own size, main_batched.py:276-287)
    The attention the reference warps with in this flow is ``blend_mask``'s image-size uint8 mask; use
    ``ops.mota_mask`` + ``ops.maps_from_attention`` when that exact quantisation is wanted -- this entry point warps
    with the token map index-upsampled on the fly (BASELINE configs[3])."""
    if len(image_sources) == 0:
        return []
    imgs = decode_jpeg_batch(image_sources, device)
    dev = imgs[0].device
    outs = ops.warp_ragged_from_tokens(tok.to(dev), imgs, out_sizes, transform=transform, exp_scale=exp_scale,
                                       exp_divisor=exp_divisor, apply_inverse=apply_inverse)
    return encode_png_batch(outs, output_paths, workers)
