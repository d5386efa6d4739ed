"""Load the UNMODIFIED reference modules from /root/reference for oracle pinning.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by
`tests/golden/make_golden.py` (run in the build container, where
/root/reference exists) and by the `-m "not gpu"` tests that re-check the
oracle against the live reference when it is present.  Nothing here is
available on the GPU box: /root/reference does not travel.

What is loaded, and how:
  * DiT: `chronoedit_diffusers/transformer_chronoedit.py` executed as-is; its
    un-vendored `diffusers==0.35.2` imports (transformer_chronoedit.py:23-32)
    resolve to `oracle/diffusers_shim/diffusers`.
  * VAE: `chronoedit/_src/tokenizers/wan2pt1.py` executed as-is with
    `sys.modules` stubs for its infra-only imports (wan2pt1.py:26-31).
  * UniPC flow-matching scheduler: `chronoedit/_src/models/fm_solvers_unipc.py` executed as-is; its diffusers
    imports (fm_solvers_unipc.py:24-28) resolve to the shim.
  * DiffSynth DiT modules (`chronoedit_diffsynth/wan_video_dit_chronoedit.py`)
    as an independent in-tree cross-check, with two import stubs.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CHRONOEDIT_REFERENCE", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusers_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "chronoedit_diffusers", "transformer_chronoedit.py"))


def _load_by_path(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name: str, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def load_reference_dit():
    """Returns the module object of the reference transformer_chronoedit.py."""
    if "_ref_transformer_chronoedit" in sys.modules:
        return sys.modules["_ref_transformer_chronoedit"]
    try:
        import diffusers  # noqa: F401  (a real install wins if it ever exists)
    except ImportError:
        sys.path.insert(0, _SHIM)
    return _load_by_path(
        "_ref_transformer_chronoedit",
        os.path.join(REFERENCE_ROOT, "chronoedit_diffusers", "transformer_chronoedit.py"),
    )


def load_reference_vae():
    """Returns the module object of the reference native Wan2.1 VAE (wan2pt1.py)."""
    if "_ref_wan2pt1" in sys.modules:
        return sys.modules["_ref_wan2pt1"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    class _Dummy:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Dummy()

        def __getattr__(self, k):
            return _Dummy()

    # infra-only imports of wan2pt1.py:26-31 that need boto3/omegaconf/...
    _stub("chronoedit._ext.imaginaire.utils.easy_io", easy_io=_Dummy())
    _stub("chronoedit._ext.imaginaire.lazy_config", LazyCall=_Dummy(), LazyDict=dict)
    return _load_by_path(
        "_ref_wan2pt1", os.path.join(REFERENCE_ROOT, "chronoedit", "_src", "tokenizers", "wan2pt1.py")
    )


def load_reference_diffsynth_dit():
    """Returns the DiffSynth DiT module (independent in-tree implementation)."""
    if "_ref_diffsynth_dit" in sys.modules:
        return sys.modules["_ref_diffsynth_dit"]
    _stub("diffsynth")
    _stub("diffsynth.models")
    _stub("diffsynth.models.utils", hash_state_dict_keys=lambda *a, **k: "")
    _stub("diffsynth.models.wan_video_camera_controller", SimpleAdapter=object)
    mod = _load_by_path(
        "_ref_diffsynth_dit",
        os.path.join(REFERENCE_ROOT, "chronoedit_diffsynth", "wan_video_dit_chronoedit.py"),
    )
    # CPU: force the plain SDPA branch of flash_attention() (wan_video_dit_chronoedit.py:43-76)
    mod.FLASH_ATTN_2_AVAILABLE = False
    mod.FLASH_ATTN_3_AVAILABLE = False
    mod.SAGE_ATTN_AVAILABLE = False
    return mod


def load_reference_unipc():
    """Returns the module object of the reference fm_solvers_unipc.py (class FlowUniPCMultistepScheduler)."""
    if "_ref_fm_solvers_unipc" in sys.modules:
        return sys.modules["_ref_fm_solvers_unipc"]
    try:
        import diffusers  # noqa: F401
    except ImportError:
        sys.path.insert(0, _SHIM)
    return _load_by_path("_ref_fm_solvers_unipc", os.path.join(REFERENCE_ROOT, "chronoedit", "_src", "models", "fm_solvers_unipc.py"))


def load_reference_pipeline():
    """Returns the module object of the UNMODIFIED chronoedit_diffusers/pipeline_chronoedit.py (class ChronoEditPipeline).
    Its diffusers imports (:25-34) resolve to the shim, `chronoedit_diffusers.transformer_chronoedit` and
    `chronoedit._ext.imaginaire.utils.log` to the reference tree itself, and the guardrail presets (:36; import chain needs nltk,
    better_profanity, retinaface ... -- only CALLED when guardrails are enabled) to an empty stub module."""
    if "_ref_pipeline_chronoedit" in sys.modules:
        return sys.modules["_ref_pipeline_chronoedit"]
    try:
        import diffusers  # noqa: F401
    except ImportError:
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    def _blocked(*a, **k):
        raise RuntimeError("guardrails are stubbed out in the oracle environment; construct the pipeline with disable_guardrails=True")

    for pkg in ("chronoedit._ext.imaginaire.auxiliary", "chronoedit._ext.imaginaire.auxiliary.guardrail",
                "chronoedit._ext.imaginaire.auxiliary.guardrail.common"):
        _stub(pkg)
    presets = _stub("chronoedit._ext.imaginaire.auxiliary.guardrail.common.presets", create_text_guardrail_runner=_blocked,
                    create_video_guardrail_runner=_blocked, run_text_guardrail=_blocked, run_video_guardrail=_blocked)
    sys.modules["chronoedit._ext.imaginaire.auxiliary.guardrail.common"].presets = presets
    return _load_by_path("_ref_pipeline_chronoedit", os.path.join(REFERENCE_ROOT, "chronoedit_diffusers", "pipeline_chronoedit.py"))
