"""ctypes binding of ``libm3b200.so`` (C ABI in ``include/m3b200.h``).

:class:`B200Session` duck-types the one method Mimic 3 calls on its
``onnxruntime.InferenceSession`` -- ``run(None, inputs)`` at reference
``mimic3_tts/voice.py:230`` -- so it can sit where ``onnx_model`` goes
(``voice.py:77,83``).  There is **no CPU fallback**: if the CUDA library or an
sm_100 GPU is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

M3_OK, M3_ERR_INVALID, M3_ERR_IO, M3_ERR_MODEL, M3_ERR_CUDA, M3_ERR_NOGPU = range(6)
FLAG_KEEP_FLOAT, FLAG_DEBUG_TENSORS, FLAG_DEVICE_IDS, FLAG_NO_HOST_COPY, FLAG_STAGE_TIMING = 1, 2, 4, 8, 16
STAGES = ("text_encoder", "duration_predictor", "durations_sync", "expand", "flow", "conv_pre", "upsample", "mrf",
          "dec_last", "post_int16")

_LIB_NAME = "libm3b200.so"
_lib = None
_lib_lock = threading.Lock()

# Every symbol include/m3b200.h declares (tests check the .so exports all of them).
API_SYMBOLS = [
    "m3_version", "m3_last_error", "m3_device_count", "m3_voice_load", "m3_voice_free",
    "m3_voice_get_info", "m3_infer", "m3_result_batch", "m3_result_sample_offsets",
    "m3_result_num_frames", "m3_result_pcm", "m3_result_audio", "m3_result_peaks",
    "m3_result_device_pcm", "m3_result_device_ms", "m3_result_kernel_launches",
    "m3_result_tensor", "m3_result_free", "m3_selftest",
    "m3_infer_ex", "m3_result_stream", "m3_wav_header",
    "m3_phoneme_table_create", "m3_phoneme_table_free", "m3_phoneme_table_load_ids", "m3_phoneme_table_load_map",
    "m3_phoneme_table_add", "m3_phoneme_table_add_map", "m3_phoneme_table_size", "m3_phoneme_table_lookup",
    "m3_phonemes_to_ids",
    "m3_voice_load_ex", "m3_voice_load_stats", "m3_weight_cache_build", "m3_weight_cache_check", "m3_sha256_file",
]


class B200EngineError(RuntimeError):
    """libm3b200 reported a failure (model, I/O, CUDA, or no sm_100 GPU)."""


class VoiceInfo(C.Structure):
    _fields_ = [
        ("num_symbols", C.c_int32), ("n_speakers", C.c_int32), ("is_multispeaker", C.c_int32),
        ("has_speaker_embedding", C.c_int32), ("sample_rate", C.c_int32), ("hop_length", C.c_int32),
        ("hidden_channels", C.c_int32), ("inter_channels", C.c_int32),
        ("noise_scale", C.c_float), ("length_scale", C.c_float), ("noise_w", C.c_float),
        ("n_params", C.c_int64), ("device", C.c_int32), ("reserved", C.c_int32),
    ]


class InferOpts(C.Structure):
    """``m3_infer_opts`` (include/m3b200.h): per-utterance settings + the PCM post chain."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("flags", C.c_uint32), ("seed", C.c_uint64),
        ("row_scales", C.POINTER(C.c_float)), ("volume", C.POINTER(C.c_double)),
        ("lead_silence", C.POINTER(C.c_int64)), ("trail_silence", C.POINTER(C.c_int64)),
        ("wav_header", C.c_int32), ("reserved", C.c_int32),
    ]


class LoadOpts(C.Structure):
    """``m3_load_opts`` (include/m3b200.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("cache_dir", C.c_char_p),
                ("expected_sha256", C.c_char_p)]


class LoadStats(C.Structure):
    """``m3_load_stats`` (include/m3b200.h): where the load time of a voice went."""
    _fields_ = [("from_cache", C.c_int32), ("cache_written", C.c_int32), ("parse_ms", C.c_double),
                ("pack_ms", C.c_double), ("cache_read_ms", C.c_double), ("hash_ms", C.c_double),
                ("upload_ms", C.c_double), ("total_ms", C.c_double), ("onnx_sha256", C.c_char * 65),
                ("cache_file", C.c_char * 512)]

    def as_dict(self) -> dict:
        return {"from_cache": bool(self.from_cache), "cache_written": bool(self.cache_written),
                "parse_ms": self.parse_ms, "pack_ms": self.pack_ms, "cache_read_ms": self.cache_read_ms,
                "hash_ms": self.hash_ms, "upload_ms": self.upload_ms, "total_ms": self.total_ms,
                "onnx_sha256": self.onnx_sha256.decode(), "cache_file": self.cache_file.decode()}


class _PhonemeOpts(C.Structure):
    """``m3_phoneme_opts`` (include/m3b200.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("blank_between", C.c_int32),
                ("n_punctuation", C.c_int32), ("bos", C.c_char_p), ("eos", C.c_char_p), ("blank", C.c_char_p),
                ("blank_word", C.c_char_p), ("punctuation_from", C.POINTER(C.c_char_p)),
                ("punctuation_to", C.POINTER(C.c_char_p)), ("separate", C.POINTER(C.c_char_p)),
                ("n_separate", C.c_int32), ("reserved", C.c_int32)]


LOAD_VERIFY_SHA256 = 1
LOAD_NO_CACHE_WRITE = 2


def library_path() -> Path:
    env = os.environ.get("M3B200_LIBRARY")
    return Path(env) if env else Path(__file__).resolve().parent / _LIB_NAME


def load_library() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) if it is missing."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not path.exists():
            raise B200EngineError(
                f"{path} not found: build it with `python -m mimic3_b200.build` "
                "(the engine has no CPU fallback)")
        lib = C.CDLL(str(path))
        vp, i32, i64p = C.c_void_p, C.c_int32, C.POINTER(C.c_int64)
        lib.m3_version.restype = C.c_char_p
        lib.m3_last_error.restype = C.c_char_p
        lib.m3_device_count.restype = i32
        lib.m3_voice_load.restype = i32
        lib.m3_voice_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
        lib.m3_voice_load_ex.restype = i32
        lib.m3_voice_load_ex.argtypes = [C.c_char_p, i32, C.POINTER(LoadOpts), C.POINTER(vp)]
        lib.m3_voice_load_stats.restype = i32
        lib.m3_voice_load_stats.argtypes = [vp, C.POINTER(LoadStats)]
        lib.m3_weight_cache_build.restype = i32
        lib.m3_weight_cache_build.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, i32]
        lib.m3_weight_cache_check.restype = i32
        lib.m3_weight_cache_check.argtypes = [C.c_char_p, C.c_char_p]
        lib.m3_sha256_file.restype = i32
        lib.m3_sha256_file.argtypes = [C.c_char_p, C.c_char_p]
        lib.m3_voice_free.restype = None
        lib.m3_voice_free.argtypes = [vp]
        lib.m3_voice_get_info.restype = i32
        lib.m3_voice_get_info.argtypes = [vp, C.POINTER(VoiceInfo)]
        lib.m3_infer.restype = i32
        lib.m3_infer.argtypes = [vp, vp, i64p, i32, i32, C.POINTER(C.c_float), i64p, C.c_uint64,
                                 C.c_uint32, C.POINTER(vp)]
        lib.m3_infer_ex.restype = i32
        lib.m3_infer_ex.argtypes = [vp, vp, i64p, i32, i32, C.POINTER(C.c_float), i64p, C.POINTER(InferOpts),
                                    C.POINTER(vp)]
        lib.m3_result_stream.restype = C.POINTER(C.c_uint8)
        lib.m3_result_stream.argtypes = [vp, i64p]
        lib.m3_wav_header.restype = i32
        lib.m3_wav_header.argtypes = [i32, C.c_int64, C.POINTER(C.c_uint8)]
        lib.m3_result_batch.restype = i32
        lib.m3_result_batch.argtypes = [vp]
        for name, rt in (("m3_result_sample_offsets", i64p), ("m3_result_num_frames", i64p),
                         ("m3_result_pcm", C.POINTER(C.c_int16)), ("m3_result_audio", C.POINTER(C.c_float)),
                         ("m3_result_peaks", C.POINTER(C.c_float)), ("m3_result_device_pcm", vp)):
            getattr(lib, name).restype = rt
            getattr(lib, name).argtypes = [vp]
        lib.m3_result_device_ms.restype = C.c_double
        lib.m3_result_device_ms.argtypes = [vp]
        lib.m3_result_kernel_launches.restype = C.c_int64
        lib.m3_result_kernel_launches.argtypes = [vp]
        lib.m3_result_tensor.restype = i32
        lib.m3_result_tensor.argtypes = [vp, C.c_char_p, C.POINTER(C.POINTER(C.c_float)), i64p, i64p]
        lib.m3_result_free.restype = None
        lib.m3_result_free.argtypes = [vp]
        lib.m3_selftest.restype = i32
        lib.m3_selftest.argtypes = [i32, C.POINTER(C.c_double)]
        _lib = lib
        return lib


def weight_cache_build(voice_path, cache_dir, expected_sha256: Optional[str] = None) -> Path:
    """One-time conversion of ``generator.onnx`` into the packed-weight blob, without a GPU (``m3_weight_cache_build``)."""
    lib = load_library()
    out = C.create_string_buffer(1024)
    rc = lib.m3_weight_cache_build(str(voice_path).encode(), str(cache_dir).encode(),
                                   expected_sha256.encode() if expected_sha256 else None, out, 1024)
    if rc != M3_OK:
        _raise(lib, rc)
    return Path(out.value.decode())


def weight_cache_check(cache_file) -> str:
    """Validates a blob completely and returns the sha256 of the generator.onnx it was made from."""
    lib = load_library()
    out = C.create_string_buffer(65)
    rc = lib.m3_weight_cache_check(str(cache_file).encode(), out)
    if rc != M3_OK:
        _raise(lib, rc)
    return out.value.decode()


def sha256_file(path) -> str:
    """``m3_sha256_file``: what ``mimic3_tts.utils.file_sha256_sum`` computes."""
    lib = load_library()
    out = C.create_string_buffer(65)
    rc = lib.m3_sha256_file(str(path).encode(), out)
    if rc != M3_OK:
        _raise(lib, rc)
    return out.value.decode()


def _raise(lib, code: int):
    msg = lib.m3_last_error().decode("utf-8", "replace")
    if code == M3_ERR_INVALID:
        raise ValueError(msg)
    if code == M3_ERR_IO:
        raise FileNotFoundError(msg)
    raise B200EngineError(f"[m3b200 error {code}] {msg}")


class InferenceResult:
    """Outputs of one engine call.  By default ``pcm`` / ``audio`` are owned copies.  With
    ``infer(..., copy=False)`` they are zero-copy views of the engine's pinned host buffers, valid
    until ``close()`` (which turns them into copies and returns the buffers to the engine's pool)."""

    __slots__ = ("pcm", "audio", "stream", "sample_offsets", "frames", "peaks", "device_ms", "launches",
                 "tensors", "device_pcm_ptr", "hop_length", "_lib", "_res", "_session")

    def detach(self):
        """Turn the views into owned copies and release the engine buffers."""
        res, self._res = getattr(self, "_res", None), None
        if res:
            self.pcm = None if self.pcm is None else self.pcm.copy()
            self.audio = None if self.audio is None else self.audio.copy()
            self.stream = None if self.stream is None else self.stream.copy()
            self._lib.m3_result_free(res)

    def close(self):
        """Release the engine buffers; zero-copy views become invalid and are dropped."""
        res, self._res = getattr(self, "_res", None), None
        if res:
            self.pcm = self.audio = self.stream = None
            self._lib.m3_result_free(res)

    def __del__(self):  # pragma: no cover
        res, self._res = getattr(self, "_res", None), None
        if res:
            try:
                self._lib.m3_result_free(res)
            except Exception:
                pass

    def utterance_pcm(self, b: int) -> np.ndarray:
        """Utterance b's own samples (silences of the post chain, if any, lie between utterances)."""
        lo = int(self.sample_offsets[b])
        return self.pcm[lo:lo + int(self.frames[b]) * self.hop_length]

    def utterance_audio(self, b: int) -> np.ndarray:
        lo = int(self.sample_offsets[b])
        return self.audio[lo:lo + int(self.frames[b]) * self.hop_length]

    def stream_bytes(self) -> bytes:
        """[WAV header if asked] + all samples with their silences, as Python bytes."""
        return self.stream.tobytes() if self.stream is not None else self.pcm.tobytes()

    @property
    def total_samples(self) -> int:
        return int(self.sample_offsets[-1])


class B200Session:
    """Stands where ``onnxruntime.InferenceSession(generator.onnx)`` stands in Mimic 3."""

    def __init__(self, path, sess_options=None, providers=None, device: Optional[int] = None,
                 cache_dir=None, expected_sha256: Optional[str] = None, verify_sha256: bool = False):
        """``cache_dir``: directory of packed-weight blobs (``*.m3w``); ``None`` = ``$M3B200_WEIGHT_CACHE``, unset =
        no cache.  ``expected_sha256``: the voice registry's ``sha256_sum`` of generator.onnx
        (``mimic3_tts/voices.json``, checked by the reference's downloader, ``download.py:108-117``): a mismatch
        raises.  ``verify_sha256`` re-hashes the file even on a cache hit."""
        self._lib = load_library()
        if device is None:
            device = int(os.environ.get("M3B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        handle = C.c_void_p()
        opts = LoadOpts()
        opts.struct_size = C.sizeof(LoadOpts)
        opts.flags = LOAD_VERIFY_SHA256 if verify_sha256 else 0
        opts.cache_dir = str(cache_dir).encode() if cache_dir else None
        opts.expected_sha256 = expected_sha256.encode() if expected_sha256 else None
        rc = self._lib.m3_voice_load_ex(str(path).encode(), int(device), C.byref(opts), C.byref(handle))
        if rc != M3_OK:
            _raise(self._lib, rc)
        self._h = handle
        st = LoadStats()
        self._lib.m3_voice_load_stats(self._h, C.byref(st))
        self.load_stats = st.as_dict()
        info = VoiceInfo()
        self._lib.m3_voice_get_info(self._h, C.byref(info))
        self.info = info
        self.device = int(info.device)

    def close(self):
        """Frees the voice (weights, contexts, pinned buffers).  Results obtained with ``copy=False`` must be closed
        or detached first: their views point into those buffers."""
        if getattr(self, "_h", None):
            self._lib.m3_voice_free(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- engine-native call ----------------------------------------------------------
    def infer(self, ids: np.ndarray, lengths: np.ndarray, scales: Sequence[float],
              sid: Optional[np.ndarray] = None, seed: int = 0, keep_float: bool = False,
              debug_tensors: Sequence[str] = (), host_copy: bool = True,
              device_ids_ptr: Optional[int] = None, stage_timing: bool = False,
              device_pcm_out=None, copy: bool = True, row_scales=None, volume=None, lead_silence=None,
              trail_silence=None, wav_header: bool = False) -> InferenceResult:
        """``device_pcm_out``: optional torch int16 CUDA tensor; the packed PCM is copied into it on the
        device (for NCCL gathers) before the engine's buffers are released.

        ``row_scales`` (batch, 3) float32, ``volume`` (batch,) float64, ``lead_silence`` / ``trail_silence``
        (batch,) int64 samples and ``wav_header`` select ``m3_infer_ex``: per-utterance settings and the
        on-device PCM post chain of ``_speak_sentence_phonemes`` / ``add_break`` (tts.py:452-465, 519-543)."""
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        batch = int(lengths.shape[0])
        extended = (row_scales is not None or volume is not None or lead_silence is not None
                    or trail_silence is not None or wav_header)
        flags = 0
        if device_ids_ptr is not None:
            ids_ptr = C.c_void_p(int(device_ids_ptr))
            t_stride = int(ids)  # caller passes the row stride in `ids`
            flags |= FLAG_DEVICE_IDS
        else:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            if ids.ndim != 2 or ids.shape[0] != batch:
                raise ValueError("ids must be int64 (batch, T) matching input_lengths")
            t_stride = int(ids.shape[1])
            ids_ptr = ids.ctypes.data_as(C.c_void_p)
        sc = (C.c_float * 3)(*[float(s) for s in scales]) if scales is not None else None
        sid_p = None
        if sid is not None:
            sid = np.ascontiguousarray(sid, dtype=np.int64)
            if sid.shape != (batch,):
                raise ValueError("sid must have shape (batch,)")
            sid_p = sid.ctypes.data_as(C.POINTER(C.c_int64))
        if keep_float:
            flags |= FLAG_KEEP_FLOAT
        if debug_tensors:
            flags |= FLAG_DEBUG_TENSORS
        if stage_timing:
            flags |= FLAG_STAGE_TIMING
            debug_tensors = tuple(debug_tensors) + tuple("ms:" + s for s in STAGES)
        if not host_copy:
            flags |= FLAG_NO_HOST_COPY
        res = C.c_void_p()
        if extended:
            keep = []  # arrays must outlive the call

            def arr(x, dtype, shape, what):
                a = np.ascontiguousarray(x, dtype=dtype)
                if a.shape != shape:
                    raise ValueError(f"{what} must have shape {shape}")
                keep.append(a)
                return a
            o = InferOpts()
            o.struct_size = C.sizeof(InferOpts)
            o.flags = flags
            o.seed = seed & (2 ** 64 - 1)
            if row_scales is not None:
                o.row_scales = arr(row_scales, np.float32, (batch, 3), "row_scales").ctypes.data_as(C.POINTER(C.c_float))
            if volume is not None:
                o.volume = arr(volume, np.float64, (batch,), "volume").ctypes.data_as(C.POINTER(C.c_double))
            if lead_silence is not None:
                o.lead_silence = arr(lead_silence, np.int64, (batch,), "lead_silence").ctypes.data_as(C.POINTER(C.c_int64))
            if trail_silence is not None:
                o.trail_silence = arr(trail_silence, np.int64, (batch,), "trail_silence").ctypes.data_as(C.POINTER(C.c_int64))
            o.wav_header = 1 if wav_header else 0
            rc = self._lib.m3_infer_ex(self._h, ids_ptr, lengths.ctypes.data_as(C.POINTER(C.c_int64)), batch,
                                       t_stride, sc, sid_p, C.byref(o), C.byref(res))
        else:
            if sc is None:
                raise ValueError("scales are required (or pass row_scales)")
            rc = self._lib.m3_infer(self._h, ids_ptr, lengths.ctypes.data_as(C.POINTER(C.c_int64)), batch,
                                    t_stride, sc, sid_p, C.c_uint64(seed & (2 ** 64 - 1)), flags, C.byref(res))
        if rc != M3_OK:
            _raise(self._lib, rc)
        out = InferenceResult()
        out._lib, out._res = self._lib, res
        out._session = self   # zero-copy views point into buffers the voice owns: keep it alive as long as the result
        out.hop_length = int(self.info.hop_length)
        out.stream = None
        off = np.ctypeslib.as_array(self._lib.m3_result_sample_offsets(res), (batch + 1,)).copy()
        out.sample_offsets = off
        out.frames = np.ctypeslib.as_array(self._lib.m3_result_num_frames(res), (batch,)).copy()
        out.peaks = np.ctypeslib.as_array(self._lib.m3_result_peaks(res), (batch,)).copy()
        total = int(off[-1])
        out.pcm = out.audio = None
        if host_copy:
            out.pcm = np.ctypeslib.as_array(self._lib.m3_result_pcm(res), (max(total, 1),))[:total]
            nbytes = C.c_int64()
            sp = self._lib.m3_result_stream(res, C.byref(nbytes))
            if nbytes.value:
                out.stream = np.ctypeslib.as_array(sp, (nbytes.value,))
            if keep_float:
                out.audio = np.ctypeslib.as_array(self._lib.m3_result_audio(res), (max(total, 1),))[:total]
        out.device_ms = float(self._lib.m3_result_device_ms(res))
        out.launches = int(self._lib.m3_result_kernel_launches(res))
        out.device_pcm_ptr = self._lib.m3_result_device_pcm(res)
        if device_pcm_out is not None and total:
            import torch

            class _View:  # zero-copy view of the engine's device buffer
                __cuda_array_interface__ = {"shape": (total,), "typestr": "<i2",
                                            "data": (int(out.device_pcm_ptr), False), "version": 2}
            device_pcm_out[:total].copy_(torch.as_tensor(_View(), device=device_pcm_out.device))
            torch.cuda.current_stream(device_pcm_out.device).synchronize()
        out.tensors = {}
        for name in debug_tensors:
            data = C.POINTER(C.c_float)()
            rows, cols = C.c_int64(), C.c_int64()
            rc = self._lib.m3_result_tensor(res, name.encode(), C.byref(data), C.byref(rows), C.byref(cols))
            if rc == M3_OK:
                n = rows.value * cols.value
                out.tensors[name] = np.ctypeslib.as_array(data, (max(n, 1),))[:n].copy().reshape(rows.value, cols.value)
        if copy or not host_copy:
            out.detach()  # owned copies (or nothing on the host): give the context back immediately
        return out

    # -- onnxruntime-compatible call (voice.py:230) ---------------------------------------
    def run(self, output_names, input_feed: Dict[str, np.ndarray], run_options=None) -> List[np.ndarray]:
        """``[float32 (B, 1, S_max)]`` like the exported graph's "output" (zero padded)."""
        for key in input_feed:
            if key not in ("input", "input_lengths", "scales", "sid"):
                raise ValueError(f"Invalid input name: {key}")
        for key in ("input", "input_lengths", "scales"):
            if key not in input_feed:
                raise ValueError(f"Missing input: {key}")
        if self.info.has_speaker_embedding and "sid" not in input_feed:
            raise ValueError("Missing input: sid")
        scales = np.asarray(input_feed["scales"], dtype=np.float32).reshape(-1)
        if scales.shape[0] != 3:
            raise ValueError("scales must have 3 elements [noise_scale, length_scale, noise_w]")
        sid = input_feed.get("sid") if self.info.has_speaker_embedding else None
        r = self.infer(np.asarray(input_feed["input"]), np.asarray(input_feed["input_lengths"]).reshape(-1),
                       scales, None if sid is None else np.asarray(sid).reshape(-1), keep_float=True,
                       seed=int.from_bytes(os.urandom(8), "little") if (scales[0] or scales[2]) else 0)
        batch = len(r.frames)
        smax = int(np.max(np.diff(r.sample_offsets)))
        out = np.zeros((batch, 1, smax), dtype=np.float32)
        for b in range(batch):
            a = r.utterance_audio(b)
            out[b, 0, : a.shape[0]] = a
        return [out]

    def get_providers(self):
        return ["B200ExecutionProvider"]
