"""attwarp_b200 -- B200-native (sm_100a) implementation of AttWarp's attention-guided warp path.

The package mirrors the reference's Python entry points for this path and nothing else:

    attwarp_b200.new_method            warp_image_by_attention, set_transform_function,
                                       save_warped_image          (AGW/new_method.py)
    attwarp_b200.checkpoint_utils      warp_from_cdf_torch, cdf_from_density, gt_marginals,
                                       upsample_pdf_right_inverse (mnfd/checkpoint_utils.py)
    attwarp_b200.model                 safe_softmax, mix_with_uniform (mnfd/model.py)
    attwarp_b200.attention_extraction  MaskHookLogger, BatchMaskHookLogger (AGW/.../llava.py)
    attwarp_b200.ops                   device-resident batched operators (torch tensors)
    attwarp_b200.sharding              image-index sharding across the GPUs of one box

All arithmetic runs in ``libattwarp_sm100.so`` (C ABI: include/attwarp.h, built by
``python -m attwarp_b200.build``).  There is no CPU fallback: importing the package is cheap,
the first call loads the library and raises if it has not been built.
"""

__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .new_method import (save_warped_image, set_transform_function,  # noqa: F401
                         warp_image_by_attention)
