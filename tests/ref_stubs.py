"""TEST INFRASTRUCTURE: import the UNMODIFIED reference package (``mimic3_tts``, ``opentts_abc``) in a sandbox that
lacks its third-party dependencies.

The reference is installed once, unmodified, into ``baseline/_ref`` (``pip install --no-deps --target baseline/_ref``,
DESIGN.md §5; git-ignored, travels to the GPU box); here (build container) ``/root/reference`` is the fallback.  Its
imports that are absent offline are replaced by stand-ins that the ids->audio path either never touches (xdgenvpy,
gruut, espeak_phonemizer, epitran, requests, tqdm) or that are small, documented pieces of test scaffolding:

* ``dataclasses_json.DataClassJsonMixin`` -- ``from_json`` / ``from_dict`` / ``to_dict`` over dataclass type hints (what
  ``TrainingConfig.load`` needs, ``config.py:325-327``);
* ``gruut_ipa.IPA`` -- the two break constants and ``graphemes`` (``config.py:171-172``, ``voice.py:713-716``);
* ``phonemes2ids`` -- the repo's restatement (``mimic3_b200/phonemes.py``), only when the real package is missing;
* ``onnxruntime`` -- ``mimic3_b200.plugin.install_as_onnxruntime`` (the product's own shim), only when missing.
"""
import dataclasses
import enum
import json
import sys
import types
import typing
from pathlib import Path
from unittest import mock

ROOT = Path(__file__).resolve().parent.parent


def reference_root():
    for p in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        if (p / "mimic3_tts" / "voice.py").is_file():
            return p
    return None


def _build(tp, value):
    """value (parsed JSON) -> instance of type hint `tp`."""
    if value is None:
        return None
    origin = typing.get_origin(tp)
    if origin is typing.Union:
        args = [a for a in typing.get_args(tp) if a is not type(None)]
        for a in args:   # dataclasses / enums first, then plain types
            if dataclasses.is_dataclass(a) and isinstance(value, dict):
                return _build(a, value)
        for a in args:
            if isinstance(a, type) and issubclass(a, enum.Enum):
                try:
                    return a(value)
                except ValueError:
                    continue
        return value
    if origin in (list, typing.List):
        (a,) = typing.get_args(tp) or (typing.Any,)
        return [_build(a, v) for v in value]
    if origin in (tuple, typing.Tuple):
        return tuple(value)
    if origin in (dict, typing.Dict):
        return dict(value)
    if dataclasses.is_dataclass(tp) and isinstance(value, dict):
        hints = typing.get_type_hints(tp)
        kw = {f.name: _build(hints[f.name], value[f.name]) for f in dataclasses.fields(tp) if f.name in value}
        return tp(**kw)
    if isinstance(tp, type) and issubclass(tp, enum.Enum):
        return tp(value)
    return value


class DataClassJsonMixin:
    @classmethod
    def from_dict(cls, d, **kw):
        return _build(cls, d)

    @classmethod
    def from_json(cls, s, **kw):
        return _build(cls, json.loads(s))

    def to_dict(self, **kw):
        return dataclasses.asdict(self)

    def to_json(self, **kw):
        return json.dumps(self.to_dict())


class _Val:
    def __init__(self, v):
        self.value = v


def _stub(name, **attrs):
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the unmodified ``mimic3_tts`` package (or None if no copy of the reference is reachable)."""
    root = reference_root()
    if root is None:
        return None
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    if "dataclasses_json" not in sys.modules:
        try:
            import dataclasses_json  # noqa: F401
        except ImportError:
            _stub("dataclasses_json", DataClassJsonMixin=DataClassJsonMixin)
    try:
        import gruut_ipa  # noqa: F401
    except ImportError:
        _stub("gruut_ipa", IPA=types.SimpleNamespace(BREAK_MINOR=_Val("|"), BREAK_MAJOR=_Val("‖"),
                                                    graphemes=lambda s: list(s)))
    try:
        import phonemes2ids  # noqa: F401
    except ImportError:
        import mimic3_b200.phonemes as ph
        sys.modules["phonemes2ids"] = ph
    try:
        import onnxruntime  # noqa: F401
    except ImportError:
        from mimic3_b200 import plugin
        plugin.install_as_onnxruntime(force=True)
    try:
        import xdgenvpy  # noqa: F401
    except ImportError:
        xdg = mock.MagicMock()
        xdg.return_value.XDG_DATA_HOME = "/nonexistent-xdg-data-home"
        xdg.return_value.XDG_DATA_DIRS = ""
        _stub("xdgenvpy", XDG=xdg)
    for name in ("espeak_phonemizer", "epitran", "gruut", "gruut.const", "gruut.utils", "gruut.text_processor",
                 "requests", "tqdm", "tqdm.auto"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name, __getattr__=lambda attr: mock.MagicMock())
    import mimic3_tts
    return mimic3_tts
