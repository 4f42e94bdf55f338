"""The on-disk formats either side of the warp: GPU JPEG decode into the BGR HWC buffers
stage 5 reads, PNG files written by the reference's own encoder.

Decode is NOT bit-exact across JPEG decoders (IDCT rounding, chroma up-sampling); the bound this test holds against
Pillow -- the reader the reference drivers use, AGW/main.py:152 -- is stated per sub-sampling mode and the measured
figures are printed.  Everything after the decode is exact: file -> warped file through ``warp_files`` equals the
oracle run on the SAME decoded pixels, and PNG files round-trip bit for bit."""

import io
import os

import numpy as np
import pytest
import torch

from gpu_util import need_gpu
from oracle import numpy_path as ON

pytestmark = pytest.mark.gpu


def _photo_like(rng, h, w):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 100 * np.sin(x / 17 + y / 23), 127 + 90 * np.cos(x / 11 - y / 29), 255 * (x + y) / (h + w)], -1)
    base += rng.normal(0, 6, base.shape)
    base[h // 3: h // 2, w // 4: w // 2] = [230, 40, 60]          # a saturated block: colour edges
    return np.clip(np.rint(base), 0, 255).astype(np.uint8)


def _jpeg_bytes(rgb, subsampling, quality=90):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, format="JPEG", quality=quality, subsampling=subsampling)
    return buf.getvalue()


@pytest.mark.parametrize("subsampling,max_lsb,mean_lsb,p99_lsb", [(0, 6, 0.7, 3), (2, None, 2.5, 16)])
def test_gpu_jpeg_decode_vs_pillow(subsampling, max_lsb, mean_lsb, p99_lsb):
    """subsampling 0 = 4:4:4 (only IDCT / colour-conversion rounding differs: measured max 4, mean 0.5 LSB),
    2 = 4:2:0 (nvJPEG replicates chroma samples, libjpeg interpolates them: measured mean 1.9 LSB, but ~100 LSB on
    the one-pixel rim of the saturated block -- no bound on the maximum is claimed for sub-sampled chroma).
    ``exact=True`` (host Pillow decode) is bit-identical in both modes."""
    need_gpu()
    from PIL import Image
    from attwarp_b200 import image_io
    rng = np.random.default_rng(8 + subsampling)
    files = [_jpeg_bytes(_photo_like(rng, h, w), subsampling) for h, w in ((336, 336), (301, 224), (480, 640))]
    files.append(_jpeg_bytes(_photo_like(rng, 200, 300)[..., 0], 0))                 # a grey JPEG
    got = image_io.decode_jpeg_batch(files)
    exact = image_io.decode_jpeg_batch(files, exact=True)
    worst, means, p99s = 0, [], []
    for buf, t, te in zip(files, got, exact):
        ref = np.array(Image.open(io.BytesIO(buf)).convert("RGB"))[..., ::-1]       # the reference's read, as BGR
        assert tuple(t.shape) == ref.shape and t.dtype == torch.uint8 and t.is_cuda and t.is_contiguous()
        assert np.array_equal(te.cpu().numpy(), ref)
        d = np.abs(t.cpu().numpy().astype(int) - ref.astype(int))
        worst, means, p99s = max(worst, int(d.max())), means + [float(d.mean())], p99s + [float(np.percentile(d, 99))]
    print(f"[jpeg decode, subsampling {subsampling}] max |diff| {worst} LSB, mean {max(means):.3f} LSB, p99 {max(p99s):.1f} LSB")
    assert max(means) <= mean_lsb and max(p99s) <= p99_lsb
    if max_lsb is not None:
        assert worst <= max_lsb


def test_png_sources_and_png_files_are_exact(tmp_path):
    need_gpu()
    import cv2
    from attwarp_b200 import image_io
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((64, 80), (97, 53))]
    srcs = []
    for k, im in enumerate(imgs):
        p = str(tmp_path / f"in{k}.png")
        cv2.imwrite(p, im)
        srcs.append(p)
    dec = image_io.decode_jpeg_batch(srcs)
    for im, t in zip(imgs, dec):
        assert np.array_equal(t.cpu().numpy(), im)
    outs = [str(tmp_path / f"out{k}.png") for k in range(2)]
    assert image_io.encode_png_batch(dec, outs) == [True, True]
    for im, p in zip(imgs, outs):
        assert np.array_equal(cv2.imread(p, cv2.IMREAD_UNCHANGED), im)


def test_warp_files_matches_oracle_on_decoded_pixels(tmp_path):
    """JPEG files in, PNG files out: equal (+-1 LSB) to the oracle warp of the pixels the GPU
    decoder produced -- the decode tolerance is the only inexact link and it is held separately above."""
    need_gpu()
    import cv2
    from attwarp_b200 import image_io
    rng = np.random.default_rng(11)
    sizes = [(224, 224), (336, 500), (301, 224)]
    files = [_jpeg_bytes(_photo_like(rng, h, w), 2) for h, w in sizes]
    tok = rng.random((3, 24, 24)) ** 3
    tok = (tok / tok.sum(axis=(1, 2), keepdims=True)).astype(np.float32)
    outs = [str(tmp_path / f"w{k}.png") for k in range(3)]
    out_sizes = [(224, 224), (500, 500), (301, 224)]
    assert image_io.warp_files(files, torch.from_numpy(tok), outs, out_sizes) == [True] * 3
    dec = image_io.decode_jpeg_batch(files)
    for k in range(3):
        full = ON.upsample_tokens_nearest(tok[k], *sizes[k])
        ref = ON.warp_image_by_attention(dec[k].cpu().numpy(), full, out_sizes[k][1], out_sizes[k][0], "identity")
        got = cv2.imread(outs[k], cv2.IMREAD_UNCHANGED)
        d = np.abs(got.astype(int) - ref.astype(int))
        assert got.shape == ref.shape and d.max() <= 1 and (d != 0).mean() <= 1e-3
