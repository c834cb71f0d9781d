"""mimic3_b200 -- B200-native (sm_100a) engine for Mimic 3's ids->waveform hot path.

Only the hot path of MycroftAI/mimic3 is implemented here: the
``onnxruntime.InferenceSession.run`` + ``audio_float_to_int16`` pair inside
``Mimic3Voice.ids_to_audio`` (reference ``mimic3_tts/voice.py:154-243``).
"""
__version__ = "0.1.0"
