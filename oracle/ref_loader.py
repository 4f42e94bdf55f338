"""Import the UNMODIFIED reference from ``/root/reference`` (build container only).

TEST / BENCH INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box; there the module
falls back to ``baseline/_ref/`` (the same files, placed by ``oracle/make_ref.py``).  Used by the
``tests/golden/make_golden*.py`` fixture generators, by CPU tests that are skipped when no tree is
present, and by the CPU arm of ``bench.py`` (``oracle/cpu_baseline.py``).  Nothing is copied: the reference files are imported from where
they lie, after stubbing the packages they import at module scope but which are not installed
(``matplotlib``, ``llava``) -- SURVEY.md Appendix A.
"""

from __future__ import annotations

import os
import sys
import types

_TRAVEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _pick_root():
    """/root/reference in the build container; on the GPU box the byte-for-byte copy of the hot-path files that
    ``oracle/make_ref.py`` placed under the git-ignored ``baseline/_ref/``."""
    env = os.environ.get("ATTWARP_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _TRAVEL):
        if os.path.isfile(os.path.join(cand, "Attention Guided Warping", "new_method.py")):
            return cand
    return "/root/reference"


REF_ROOT = _pick_root()
_AGW = os.path.join(REF_ROOT, "Attention Guided Warping")
_MNFD = os.path.join(REF_ROOT, "model", "marginalnet_full_dataset")


def available() -> bool:
    return os.path.isfile(os.path.join(_AGW, "new_method.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _stub_matplotlib():
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("mpl_toolkits")
    _stub("mpl_toolkits.axes_grid1")
    _stub("mpl_toolkits.axes_grid1.inset_locator", inset_axes=None, mark_inset=None)


def _stub_llava():
    _stub("llava")
    _stub("llava.constants", IMAGE_TOKEN_INDEX=-200, DEFAULT_IMAGE_TOKEN="<image>",
          DEFAULT_IM_START_TOKEN="<im_start>", DEFAULT_IM_END_TOKEN="<im_end>",
          IMAGE_PLACEHOLDER="<image-placeholder>")
    _stub("llava.conversation", conv_templates={}, SeparatorStyle=None)
    _stub("llava.model")
    _stub("llava.model.builder", load_pretrained_model=None)
    _stub("llava.utils", disable_torch_init=lambda: None)
    _stub("llava.mm_utils", process_images=None, tokenizer_image_token=None,
          get_model_name_from_path=None, KeywordsStoppingCriteria=None)
    try:
        import transformers.generation.stopping_criteria as sc
        if not hasattr(sc, "MaxNewTokensCriteria"):
            sc.MaxNewTokensCriteria = object
    except Exception:
        pass


_cache = {}


def new_method():
    """``Attention Guided Warping/new_method.py`` as a module."""
    if "new_method" not in _cache:
        if _AGW not in sys.path:
            sys.path.insert(0, _AGW)
        _stub_llava()
        import importlib
        _cache["new_method"] = importlib.import_module("new_method")
    return _cache["new_method"]


def checkpoint_utils():
    """``model/marginalnet_full_dataset/checkpoint_utils.py`` as a module."""
    if "checkpoint_utils" not in _cache:
        if _MNFD not in sys.path:
            sys.path.insert(0, _MNFD)
        _stub_matplotlib()
        _stub("tqdm", tqdm=lambda x, **k: x) if "tqdm" not in sys.modules else None
        import importlib
        _cache["checkpoint_utils"] = importlib.import_module("checkpoint_utils")
    return _cache["checkpoint_utils"]


def marginalnet_model():
    """``model/marginalnet_full_dataset/model.py`` as a module (named ``model``)."""
    if "model" not in _cache:
        import importlib.util
        spec = importlib.util.spec_from_file_location("attwarp_ref_mnfd_model",
                                                      os.path.join(_MNFD, "model.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache["model"] = mod
    return _cache["model"]


def llava_hooks():
    """``attention_extraction/llava.py`` as a module (MaskHookLogger, BatchMaskHookLogger,
    revise_mask, blend_mask)."""
    if "llava_hooks" not in _cache:
        if _AGW not in sys.path:
            sys.path.insert(0, _AGW)
        _stub_llava()
        import importlib
        pkg = importlib.import_module("attention_extraction.llava")
        _cache["llava_hooks"] = pkg
    return _cache["llava_hooks"]
