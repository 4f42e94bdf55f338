"""Mirror of the hook-logger attention reducers of
``Attention Guided Warping/attention_extraction/llava.py``.

* ``MaskHookLogger``       (llava.py:37-153)  -- single sample per generate call
* ``BatchMaskHookLogger``  (llava.py:338-448) -- per-sample image-token ranges
* ``revise_mask`` / ``blend_mask`` (llava.py:223-270) -- mask post-processing and the image-size
  uint8 mask the drivers warp with (the JET overlay half of ``blend_mask`` is visualisation and stays on
  OpenCV like in the reference)

Same constructor / method names.  ``_process_attention`` hands the live
``[B, Hh, q, kv]`` attention tensor (fp16/bf16/fp32, on the GPU, no copy) to the stage-1 CUDA
kernel, which slices the last query row at the per-sample token offset, renormalises per head,
averages over heads and adds the step into a running device-side sum; ``finalize`` /
``finalize_batch`` divide by the number of steps.  The reference's Python loop over the batch
(llava.py:388-395) and its list/stack of per-step tensors are gone.  The input is upcast to
fp32 in the kernel (the reference divides in the tensor's dtype).
"""

from __future__ import annotations

import ctypes as _C
import os as _os

import torch

from . import _lib, ops

# ATTWARP_HOOK_FAST=0 sends every hooked step through ops.aggregate_attention (A/B of the host-side cost)
_HOOK_FAST = _os.environ.get("ATTWARP_HOOK_FAST", "1") != "0"


class _RunningAttention:
    """Device-side running sum of per-step head means."""

    def __init__(self):
        self.sum = None
        self.steps = 0
        self._starts_key = None          # (device, tuple of starts) of the cached offset tensor
        self._starts = None
        self._fast = None                # (layout key, starts, ends, prepared C call) of the previous step

    def _start_tensor(self, starts, device):
        """int32 device tensor of the per-sample offsets, rebuilt only when they change: generate() calls
        the hook once per decoding step with the same ranges, and a host->device copy per step was most of
        the hooked step's cost."""
        key = (device, tuple(starts))
        if key != self._starts_key:
            self._starts = torch.tensor(key[1], dtype=torch.int32, device=device)
            self._starts_key = key
        return self._starts

    def add_step(self, attn_weights: torch.Tensor, starts, ends):
        # generate() calls the hook once per layer and decoding step with the same layout and token ranges: after the
        # first step of a layout the validated arguments are kept and a step is one C call: the hook is bound by its
        # host side (slicing, checks, tensor conversions, device guard), 42-45 -> 24-25 us per hooked step
        # (profiles/r08d_hook_overhead.txt)
        fast = self._fast
        if fast is not None and type(starts) is list and type(ends) is list:
            shape = attn_weights.shape           # [B, Hh, q, kv]: q and kv change from step to step (KV cache)
            if (fast[0] == (shape[0], shape[1], attn_weights.dtype, attn_weights.device) and shape[3] >= fast[4]
                    and fast[1] == starts and fast[2] == ends and attn_weights.stride(3) == 1):
                fast[3](attn_weights)
                self.steps += 1
                return
        self._fast = None
        B, Hh, q, kv = attn_weights.shape
        starts = [int(s) for s in starts]
        ends = [min(int(e), kv) for e in ends]
        lens = sorted({e - s for s, e in zip(starts, ends)})
        if len(lens) != 1:
            raise ValueError(f"per-sample image-token spans differ in length: {lens}")
        T = lens[0]
        if T <= 0:
            raise ValueError(f"empty image-token span (start {starts[0]}, end {ends[0]}, kv length {kv})")
        if len(starts) != B:
            raise ValueError(f"{len(starts)} image-token ranges for a batch of {B}")
        if min(starts) < 0:
            raise ValueError(f"negative image-token start in {starts}")
        rows = attn_weights[:, :, -1, :].unsqueeze(1)          # [B, 1, Hh, kv] view, no copy
        if rows.stride(3) != 1:
            rows = rows.contiguous()
        if self.sum is None:
            self.sum = torch.zeros(B, T, dtype=torch.float32, device=rows.device)
        elif tuple(self.sum.shape) != (B, T) or self.sum.device != rows.device:
            # the reference fails in torch.stack / torch.cat when steps disagree (llava.py:131, 409)
            raise ValueError(f"attention step of shape (B={B}, T={T}) on {rows.device} does not match the running "
                             f"sum {tuple(self.sum.shape)} on {self.sum.device}; call reinit() between batches")
        st = self._start_tensor(starts, rows.device)
        ops.aggregate_attention(rows, tok_start=st, num_tokens=T, out=self.sum, accumulate=True)
        self.steps += 1
        if _HOOK_FAST and rows.data_ptr() == attn_weights.data_ptr() + (q - 1) * attn_weights.stride(2) * attn_weights.element_size():
            # (the ranges as normalised above: a later step matches only if its raw ranges need no clipping)
            self._fast = ((B, Hh, attn_weights.dtype, attn_weights.device), list(starts), list(ends),
                          self._prepare(attn_weights, st, T), max(ends))

    def _prepare(self, attn_weights, st, T):
        """The C call of one more step on a ``[B, Hh, q, kv]`` tensor of this batch size, head count, dtype and
        device (its data pointer, strides and last query row, the workspace of the current stream and the stream are
        read per call; everything else is fixed)."""
        lib = _lib.load()
        fn = lib.attwarp_aggregate_attention
        B, Hh = attn_weights.shape[0], attn_weights.shape[1]
        dev = attn_weights.device
        esize = attn_weights.element_size()
        dtype_id = _lib.TORCH_DTYPE_IDS[attn_weights.dtype]
        wsb = lib.attwarp_aggregate_workspace_bytes(B, 1, Hh, T)
        st_ptr, sum_ptr = _C.c_void_p(st.data_ptr()), _C.c_void_p(self.sum.data_ptr())
        keep = (st, self.sum)                                                         # the pointers above stay valid
        dev_index = dev.index

        def call(t):
            s0, s1, s2, _ = t.stride()
            row_off = (t.shape[2] - 1) * s2 * esize                                   # the last query row
            ws = ops._workspace(wsb, dev)
            stream = _C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if torch.cuda.current_device() == dev_index:
                rc = fn(_C.c_void_p(t.data_ptr() + row_off), dtype_id, B, 1, Hh, T, s0, 0, s1, st_ptr, 1e-12,
                        _C.c_void_p(ws.data_ptr()), ws.numel(), sum_ptr, 1, 1.0, stream)
            else:
                with torch.cuda.device(dev):
                    rc = fn(_C.c_void_p(t.data_ptr() + row_off), dtype_id, B, 1, Hh, T, s0, 0, s1, st_ptr, 1e-12,
                            _C.c_void_p(ws.data_ptr()), ws.numel(), sum_ptr, 1, 1.0, stream)
            _lib.check(rc)
            return keep
        return call

    def mean(self):
        return self.sum / float(self.steps)


class MaskHookLogger(object):
    """Captures last-token -> image-token attention of one decoder layer during generation."""

    def __init__(self, model, device, layer_index=20):
        self.device = device
        self.model = model
        self.layer_index = layer_index
        self.hook_handle = None
        self.image_token_start = None
        self.image_token_end = None
        self.num_image_tokens = 576  # 24x24 patches for LLaVA-1.5
        self._acc = _RunningAttention()

    @property
    def attns(self):
        """Number-of-steps view kept for callers that test ``len(logger.attns)``."""
        return [None] * self._acc.steps

    @torch.no_grad()
    def _attention_hook(self, module, input, output):
        if isinstance(output, tuple) and len(output) >= 2:
            attn_weights = output[1]
            if attn_weights is not None and isinstance(attn_weights, torch.Tensor):
                if len(attn_weights.shape) == 4:
                    self._process_attention(attn_weights)

    @torch.no_grad()
    def _process_attention(self, attn_weights):
        kv = attn_weights.shape[-1]
        if self.image_token_start is None or self.image_token_end is None:
            st, ed = 1, min(1 + self.num_image_tokens, kv)            # llava.py:99-102
        else:
            st, ed = self.image_token_start, min(self.image_token_end, kv)
        B = attn_weights.shape[0]
        self._acc.add_step(attn_weights.detach(), [st] * B, [ed] * B)

    def set_image_token_range(self, start, end):
        self.image_token_start = start
        self.image_token_end = end

    @torch.no_grad()
    def finalize(self):
        """[num_image_tokens] mean over steps (and, like llava.py:131-132, over the batch)."""
        if self._acc.steps == 0:
            return torch.ones(self.num_image_tokens, device=self.device) / self.num_image_tokens
        return self._acc.mean().mean(dim=0).to(self.device)

    def reinit(self):
        self._acc = _RunningAttention()
        self.image_token_start = None
        self.image_token_end = None

    def register_hook(self):
        if self.hook_handle is not None:
            self.hook_handle.remove()
        attn_layer = self.model.model.layers[self.layer_index].self_attn
        self.hook_handle = attn_layer.register_forward_hook(self._attention_hook)

    def remove_hook(self):
        if self.hook_handle is not None:
            self.hook_handle.remove()
            self.hook_handle = None


class BatchMaskHookLogger(object):
    """Batched variant with per-sample image-token ranges (left-padding offsets)."""

    def __init__(self, model, device, layer_index=20):
        self.device = device
        self.model = model
        self.layer_index = layer_index
        self.hook_handle = None
        self.num_image_tokens = 576
        self.image_token_starts = None
        self.image_token_ends = None
        self.batch_size = 0
        self._acc = _RunningAttention()
        self._original_forward = None

    @property
    def step_attentions(self):
        return [None] * self._acc.steps

    def set_batch_image_token_ranges(self, starts, ends):
        assert len(starts) == len(ends)
        self.image_token_starts = starts
        self.image_token_ends = ends
        self.batch_size = len(starts)

    @torch.no_grad()
    def _attention_hook(self, module, input, output):
        if not isinstance(output, tuple) or len(output) < 2:
            return
        attn_weights = output[1]
        if attn_weights is None or not isinstance(attn_weights, torch.Tensor):
            return
        if len(attn_weights.shape) != 4:
            return
        self._process_attention(attn_weights)

    @torch.no_grad()
    def _process_attention(self, attn_weights):
        bsz = attn_weights.shape[0]
        self._acc.add_step(attn_weights.detach(), self.image_token_starts[:bsz],
                           self.image_token_ends[:bsz])

    @torch.no_grad()
    def finalize_batch(self):
        """List of [24, 24] maps, one per sample (llava.py:401-411)."""
        if self._acc.steps == 0:
            return [torch.ones(self.num_image_tokens, device=self.device) / self.num_image_tokens
                    for _ in range(self.batch_size)]
        avg = self._acc.mean()
        return [avg[i].view(24, 24) for i in range(self.batch_size)]

    def reinit(self):
        self._acc = _RunningAttention()
        self.image_token_starts = None
        self.image_token_ends = None
        self.batch_size = 0

    def register_hook_and_patch(self):
        """Hook the layer and force ``output_attentions=True`` on it only (llava.py:422-438)."""
        if self.hook_handle is not None:
            self.hook_handle.remove()
        attn_layer = self.model.model.layers[self.layer_index].self_attn
        self.hook_handle = attn_layer.register_forward_hook(self._attention_hook)
        self._original_forward = attn_layer.forward
        original = self._original_forward

        def forward_with_attentions(*args, **kwargs):
            kwargs["output_attentions"] = True
            return original(*args, **kwargs)

        attn_layer.forward = forward_with_attentions

    def remove_hook_and_unpatch(self):
        if self.hook_handle is not None:
            self.hook_handle.remove()
            self.hook_handle = None
        if self._original_forward is not None:
            self.model.model.layers[self.layer_index].self_attn.forward = self._original_forward
            self._original_forward = None


def hook_logger(model, device, layer_index=20):
    """Create and register a ``MaskHookLogger`` (llava.py:156-187): turns ``output_attentions`` on in the
    model config, hooks the layer, and leaves the logger on ``model.hooklogger``."""
    prs = MaskHookLogger(model, device, layer_index)
    original_output_attentions = getattr(model.config, "output_attentions", False)
    model.config.output_attentions = True
    prs.register_hook()
    model.hooklogger = prs
    model._original_output_attentions = original_output_attentions
    return prs


def batch_hook_logger(model, device, layer_index=20):
    """Create and register a ``BatchMaskHookLogger`` (llava.py:451-462): only the target layer is patched
    to produce attention weights; ``model.config.output_attentions`` is switched off."""
    prs = BatchMaskHookLogger(model, device, layer_index)
    model.config.output_attentions = False
    prs.register_hook_and_patch()
    model.batch_hooklogger = prs
    return prs


# ----------------------------------------------------------------------------------------------
# mask post-processing (llava.py:207-270)
# ----------------------------------------------------------------------------------------------
def revise_mask(patch_mask: torch.Tensor, kernel_size: int = 3, enhance_coe: int = 10) -> torch.Tensor:
    """[gh, gw] token map -> [gh, gw] smoothed sigmoid mask (llava.py:223-238), on the mask's device."""
    assert kernel_size % 2 == 1
    dev = patch_mask.device if patch_mask.is_cuda else torch.device("cuda", torch.cuda.current_device())
    out = ops.revise_mask(patch_mask.detach().to(dev).float()[None], kernel_size, enhance_coe)[0]
    return out.to(patch_mask.device)


def blend_mask(image_path_or_pil_image, mask, enhance_coe, kernel_size, interpolate_method, grayscale):
    """Same signature and return value as llava.py:240-270: (overlay PIL image, mode-'L' PIL mask at the
    image size).  The mask is computed on the GPU (bit-identical to Pillow's LANCZOS resize); only LANCZOS
    -- what the drivers configure -- is implemented there."""
    import cv2
    import numpy as np
    from PIL import Image
    if isinstance(image_path_or_pil_image, str):
        image = Image.open(image_path_or_pil_image)
    elif isinstance(image_path_or_pil_image, Image.Image):
        image = image_path_or_pil_image
    else:
        raise NotImplementedError
    if interpolate_method not in (Image.LANCZOS, "LANCZOS"):
        raise NotImplementedError("blend_mask: only Image.LANCZOS is implemented on the device")
    dev = mask.device if mask.is_cuda else torch.device("cuda", torch.cuda.current_device())
    W, H = image.size
    m = ops.mota_mask(mask.detach().to(dev).float()[None], (H, W), kernel_size, enhance_coe)[0]
    mask_img = Image.fromarray(m.cpu().numpy(), mode="L")
    mask_np = np.array(mask_img).astype(np.float32)
    mask_norm = cv2.normalize(mask_np, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    heatmap_bgr = cv2.applyColorMap(mask_norm, cv2.COLORMAP_JET)
    if isinstance(image_path_or_pil_image, str):
        orig_bgr = cv2.imread(image_path_or_pil_image)
    else:
        orig_bgr = cv2.cvtColor(np.array(image.convert("RGB")), cv2.COLOR_RGB2BGR)
    alpha = grayscale if (isinstance(grayscale, (int, float)) and 0 < grayscale <= 1) else 0.5
    overlay_bgr = cv2.addWeighted(orig_bgr, 1 - alpha, heatmap_bgr, alpha, 0)
    return Image.fromarray(cv2.cvtColor(overlay_bgr, cv2.COLOR_BGR2RGB)), mask_img
