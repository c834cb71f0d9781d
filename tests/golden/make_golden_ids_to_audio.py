"""Golden vectors for the host prologue of Mimic3Voice.ids_to_audio (mimic3_tts/voice.py:154-243), produced BY THE
REFERENCE ITSELF: the unmodified reference module is imported from /root/reference with its missing third-party
imports stubbed (none of them is touched by ids_to_audio), a recording fake stands where the onnxruntime session
goes, and the real method is called over a grid of speaker / scale / rate arguments.  What the fake session receives
(names, dtypes, shapes, values of `input`, `input_lengths`, `scales`, `sid`) is the golden (the int16 conversion of
the result is pinned separately by tests/golden/int16_reference.npz).
Output: tests/golden/ids_to_audio_inputs.json      python tests/golden/make_golden_ids_to_audio.py
"""
import itertools
import json
import sys
import types
from pathlib import Path
from types import SimpleNamespace
from unittest import mock

import numpy as np

sys.path.insert(0, "/root/reference")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Val:
    def __init__(self, v):
        self.value = v


stub("xdgenvpy", XDG=mock.MagicMock())
stub("gruut_ipa", IPA=types.SimpleNamespace(BREAK_MINOR=_Val("|"), BREAK_MAJOR=_Val("‖"), graphemes=lambda s: list(s)))
stub("dataclasses_json", DataClassJsonMixin=type("DataClassJsonMixin", (), {}))
for name in ("phonemes2ids", "onnxruntime", "espeak_phonemizer", "epitran", "gruut", "gruut.const", "gruut.utils",
             "gruut.text_processor", "requests", "tqdm", "tqdm.auto"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            stub(name, __getattr__=lambda attr: mock.MagicMock())

import mimic3_tts.voice as ref  # noqa: E402  (the unmodified reference)


class RefVoice(ref.Mimic3Voice):
    def text_to_phonemes(self, text, text_language=None):  # abstract in the reference; not on this path
        raise NotImplementedError


class FakeOrt:
    def __init__(self):
        self.calls = []

    def run(self, output_names, inputs):
        assert output_names is None
        self.calls.append({k: dict(dtype=str(v.dtype), shape=list(v.shape), values=v.reshape(-1).tolist())
                           for k, v in inputs.items()})
        n = int(inputs["input"].shape[1]) * 8
        wave = (np.sin(np.arange(n, dtype=np.float32) * 0.37) * 0.25).astype(np.float32)
        return [wave.reshape(1, 1, n)]


def main():
    cases = []
    speaker_maps = {"none": None, "names": {"p239": 2, "alias": 1, "7": 4}}
    for multi, map_key in itertools.product((True, False), speaker_maps):
        config = SimpleNamespace(inference=SimpleNamespace(length_scale=1.2, noise_scale=0.5, noise_w=0.7),
                                 audio=SimpleNamespace(sample_rate=22050), is_multispeaker=multi)
        for speaker in (None, "p239", "alias", "1", "7", "nobody", 2, 0):
            for length_scale, noise_scale, noise_w, rate in ((None, None, None, 1.0), (0.9, 0.0, 0.0, 1.0),
                                                             (None, 0.333, None, 2.0), (1.5, None, 1.0, 0.5),
                                                             (None, None, None, 0.0), (2.0, 0.1, 0.2, -1.0)):
                fake = FakeOrt()
                v = RefVoice(config=config, onnx_model=fake, phoneme_to_id={}, phoneme_map=None,
                             speaker_map=speaker_maps[map_key])
                ids = [3, 4, 5, 6, 17][: 2 + (len(cases) % 4)]
                audio = v.ids_to_audio(ids, speaker=speaker, length_scale=length_scale, noise_scale=noise_scale,
                                       noise_w=noise_w, rate=rate)
                cases.append(dict(multispeaker=multi, speaker_map=speaker_maps[map_key], ids=ids, speaker=speaker,
                                  length_scale=length_scale, noise_scale=noise_scale, noise_w=noise_w, rate=rate,
                                  inputs=fake.calls[0], audio_dtype=str(audio.dtype), audio_len=int(audio.shape[0])))
    path = Path(__file__).resolve().parent / "ids_to_audio_inputs.json"
    path.write_text(json.dumps(dict(source="mimic3_tts/voice.py:154-243 executed with stubbed third-party imports",
                                    defaults=dict(length_scale=1.2, noise_scale=0.5, noise_w=0.7), cases=cases),
                               separators=(",", ":")))
    print(f"{len(cases)} calls -> {path} ({path.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
