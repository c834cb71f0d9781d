"""Selecting the B200 engine from an UNMODIFIED Mimic 3 install.

The reference creates its engine in one place -- ``Mimic3Voice._load_model``
(``mimic3_tts/voice.py:378-407``): ``onnxruntime.InferenceSession(str(generator_path),
sess_options=..., providers=...)`` -- and calls it in one place, ``voice.py:230``.  Two ways to put
``libm3b200`` there without editing the reference:

* :func:`patch_reference` wraps ``Mimic3Voice._load_model``: a ``providers`` list that names
  ``"B200ExecutionProvider"`` (or ``M3B200_PROVIDER=1`` in the environment, which also captures
  ``Mimic3Settings.use_cuda`` -> ``["CUDAExecutionProvider"]``, ``tts.py:590-593``) returns a
  :class:`~mimic3_b200.engine.B200Session`; anything else falls through to onnxruntime as before.
* :func:`install_as_onnxruntime` registers a module named ``onnxruntime`` that exposes exactly what
  ``mimic3_tts`` touches (``InferenceSession``, ``SessionOptions``, ``GraphOptimizationLevel``,
  ``get_available_providers``) for hosts that have no onnxruntime wheel at all.

Either way ``voice.py:230-231`` runs as written: ``B200Session.run(None, inputs)`` returns
``[float32 (B, 1, S)]`` and the reference keeps doing ``audio_float_to_int16`` itself.  Nothing here falls
back to a CPU path: a missing CUDA library or GPU raises from the session constructor.
"""
from __future__ import annotations

import os
import sys
import types
import typing

from .engine import B200Session

B200_PROVIDER = "B200ExecutionProvider"


def _names(providers) -> typing.List[str]:
    return [p if isinstance(p, str) else p[0] for p in (providers or [])]


def wants_b200(providers) -> bool:
    env = os.environ.get("M3B200_PROVIDER", "")
    return B200_PROVIDER in _names(providers) or (env not in ("", "0"))


class SessionOptions:
    """The two attributes ``_load_model`` sets (``voice.py:392-401``); the engine is deterministic for fixed seeds and
    noise scales 0, so ``use_deterministic_compute`` needs no action."""

    def __init__(self):
        self.graph_optimization_level = None
        self.use_deterministic_compute = False
        self.intra_op_num_threads = 0
        self.inter_op_num_threads = 0


class GraphOptimizationLevel:
    ORT_DISABLE_ALL, ORT_ENABLE_BASIC, ORT_ENABLE_EXTENDED, ORT_ENABLE_ALL = 0, 1, 2, 99


class InferenceSession(B200Session):
    """``onnxruntime.InferenceSession(path, sess_options=None, providers=None)`` signature over ``m3_voice_load``."""

    def __init__(self, path_or_bytes, sess_options=None, providers=None, provider_options=None, **kwargs):
        if not isinstance(path_or_bytes, (str, os.PathLike)):
            raise TypeError("B200 InferenceSession loads a voice from its generator.onnx PATH (config.json beside it)")
        super().__init__(os.fspath(path_or_bytes), sess_options=sess_options, providers=providers)


def get_available_providers() -> typing.List[str]:
    return [B200_PROVIDER]


def get_device() -> str:
    return "GPU"


def install_as_onnxruntime(force: bool = False) -> types.ModuleType:
    """Make ``import onnxruntime`` resolve to this shim (only when the real package is absent, unless ``force``)."""
    if not force:
        try:
            import importlib.util
            if importlib.util.find_spec("onnxruntime") is not None and "onnxruntime" not in sys.modules:
                raise RuntimeError("a real onnxruntime is installed: use patch_reference() or pass force=True")
        except (ImportError, ValueError):
            pass
    import importlib.machinery
    m = types.ModuleType("onnxruntime")
    m.__spec__ = importlib.machinery.ModuleSpec("onnxruntime", loader=None)   # importlib.util.find_spec() needs one
    m.__dict__.update(InferenceSession=InferenceSession, SessionOptions=SessionOptions,
                      GraphOptimizationLevel=GraphOptimizationLevel, get_available_providers=get_available_providers,
                      get_device=get_device, __version__="0+m3b200", __m3b200_shim__=True)
    sys.modules["onnxruntime"] = m
    return m


def patch_reference(voice_module=None):
    """Wrap ``mimic3_tts.voice.Mimic3Voice._load_model`` (idempotent).  Returns the voice module."""
    if voice_module is None:
        import mimic3_tts.voice as voice_module  # the unmodified reference
    cls = voice_module.Mimic3Voice
    if getattr(cls._load_model, "__m3b200_patched__", False):
        return voice_module
    original = cls._load_model

    def _load_model(generator_path, session_options=None, providers=None, use_deterministic_compute=False):
        if wants_b200(providers):
            return B200Session(str(generator_path), sess_options=session_options, providers=providers)
        return original(generator_path, session_options=session_options, providers=providers,
                        use_deterministic_compute=use_deterministic_compute)

    _load_model.__m3b200_patched__ = True
    _load_model.__wrapped__ = original
    cls._load_model = staticmethod(_load_model)
    return voice_module
