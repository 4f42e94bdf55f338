"""CPU oracle for the AttWarp attention-guided warp hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``attwarp_b200/`` imports this package.
It may be imported from exactly three places: ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs -- and there only as the
checker or as the timed CPU baseline, never as a product code path.

Contents
--------
``numpy_path``   float64 NumPy restatement of ``warp_image_by_attention``
                 (reference: ``Attention Guided Warping/new_method.py:133-283``),
                 including a first-principles restatement of ``np.interp`` and of
                 OpenCV's ``cv2.remap(INTER_LINEAR, BORDER_REPLICATE)`` fixed-point
                 bilinear kernel (third-party; OpenCV is unpinned by the reference,
                 4.13.0 in this image).
``torch_path``   NumPy restatement of the torch-side helpers
                 (reference: ``model/marginalnet_full_dataset/checkpoint_utils.py:17-204``,
                 ``model/marginalnet_full_dataset/model.py:8-14,98-101``).
``aggregate``    restatement of the hook-logger reducers
                 (reference: ``Attention Guided Warping/attention_extraction/llava.py:94-132,385-411``).
``mask_path``    restatement of ``revise_mask`` and the mask half of ``blend_mask``
                 (reference: ``Attention Guided Warping/attention_extraction/llava.py:207-256``),
                 including Pillow's 8-bit two-pass LANCZOS resampler (third-party, unpinned; 12.2.0 here).
``ref_loader``   imports the *unmodified* reference from ``/root/reference`` (only exists
                 in the build container) -- used to generate ``tests/golden/*.npz`` and
                 to pin the restatements.

Parity pinning: the reference ships no tests / golden vectors, so the
restatements are pinned against outputs of the reference itself, executed in the build
container by ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``;
``tests/test_oracle_vs_golden.py`` replays them.
"""

from . import numpy_path, torch_path, aggregate, mask_path  # noqa: F401
