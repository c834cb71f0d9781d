"""PCM post chain: oracle pinned against the real audioop / wave modules (CPU);
device path against the oracle through m3_infer_ex (GPU)."""
import ctypes
import warnings

import numpy as np
import pytest

from oracle import post_chain as pc


def test_oracle_volume_matches_real_audioop():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        audioop = pytest.importorskip("audioop")
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32768, size=20000).astype(np.int16)
    pcm[:6] = [-32768, -32767, 32767, 0, 1, -1]
    for factor in (0.0, 0.5, 0.3333, 1.0, 1.5, 2.0, 50 / 100.0, 150 / 100.0, 0.999999, 7.25, 1e-9):
        want = np.frombuffer(audioop.mul(pcm.tobytes(), 2, factor), dtype=np.int16)
        np.testing.assert_array_equal(pc.audioop_mul_int16(pcm, factor), want, err_msg=str(factor))


def test_wav_header_helper_matches_wave_module(built_library):
    from mimic3_b200.engine import load_library
    lib = load_library()
    for sr, n in ((22050, 0), (22050, 1), (16000, 12345), (48000, 10_000_000)):
        buf = (ctypes.c_uint8 * 44)()
        assert lib.m3_wav_header(sr, n, buf) == 0
        want = pc.wav_bytes(np.zeros(min(n, 4), dtype=np.int16), sr)[:44]
        got = bytes(buf)
        if n <= 4:
            assert got == want
        else:  # only the two size fields differ from the short reference file
            assert got[:4] == want[:4] and got[8:40] == want[8:40]
            assert int.from_bytes(got[40:44], "little") == 2 * n
            assert int.from_bytes(got[4:8], "little") == 36 + 2 * n
    assert lib.m3_wav_header(0, 10, (ctypes.c_uint8 * 44)()) != 0
    assert lib.m3_wav_header(22050, -1, (ctypes.c_uint8 * 44)()) != 0


def test_break_samples():
    assert pc.break_samples(500, 22050) == 11025
    assert pc.break_samples(1, 22050) == 22
    assert pc.break_samples(0, 22050) == 0


@pytest.mark.gpu
def test_device_post_chain_matches_host_chain(built_library, voices):
    """Volume (audioop.mul), silences (add_break) and WAV framing done by the engine == the reference's host
    passes over the same int16 audio, byte for byte."""
    import io
    import wave
    from mimic3_b200.engine import B200Session
    sess = B200Session(str(voices("tiny_ms")))
    rng = np.random.default_rng(4)
    lens = np.array([12, 30, 1, 7], dtype=np.int64)
    ids = np.zeros((4, 30), dtype=np.int64)
    for b, L in enumerate(lens):
        ids[b, :L] = rng.integers(4, 20, size=L)
    sid = np.array([0, 1, 2, 1])
    scales = (0.0, 1.0, 0.0)
    sr = sess.info.sample_rate
    plain = sess.infer(ids, lens, scales, sid)
    utts = [plain.utterance_pcm(b).copy() for b in range(4)]
    volume = np.array([1.0, 0.35, 2.5, 0.0])
    lead = np.array([pc.break_samples(10, sr), 0, 333, 0])
    trail = np.array([0, pc.break_samples(250, sr), 0, 17])
    for kw in (dict(volume=volume), dict(lead_silence=lead), dict(trail_silence=trail),
               dict(volume=volume, lead_silence=lead, trail_silence=trail),
               dict(wav_header=True), dict(volume=volume, lead_silence=lead, trail_silence=trail, wav_header=True)):
        r = sess.infer(ids, lens, scales, sid, **kw)
        want = pc.assemble_stream(utts, sr, kw.get("volume"), kw.get("lead_silence"), kw.get("trail_silence"),
                                  kw.get("wav_header", False))
        assert r.stream_bytes() == want, kw
        np.testing.assert_array_equal(r.frames, plain.frames)
        for b in range(4):  # offsets point at each utterance's own samples
            w = utts[b] if "volume" not in kw else pc.audioop_mul_int16(utts[b], volume[b])
            np.testing.assert_array_equal(r.utterance_pcm(b), w)
        if kw.get("wav_header"):
            with wave.open(io.BytesIO(r.stream_bytes()), "rb") as f:
                assert (f.getframerate(), f.getsampwidth(), f.getnchannels()) == (sr, 2, 1)
                assert f.getnframes() == int(r.sample_offsets[-1])
    with pytest.raises(ValueError, match="KEEP_FLOAT"):
        sess.infer(ids, lens, scales, sid, volume=volume, keep_float=True)
    with pytest.raises(ValueError, match="silence"):
        sess.infer(ids, lens, scales, sid, lead_silence=np.array([0, -1, 0, 0]))
    with pytest.raises(ValueError, match="volume"):
        sess.infer(ids, lens, scales, sid, volume=np.array([1.0, np.nan, 1.0, 1.0]))
    sess.close()


@pytest.mark.gpu
def test_per_utterance_scales_equal_uniform_runs(built_library, voices):
    """row_scales: row b of a mixed-settings batch == row b of a batch run entirely with b's settings
    (same Philox row index), i.e. sentences with different Mimic3Settings can share one engine call."""
    from mimic3_b200.engine import B200Session
    sess = B200Session(str(voices("low_ms")))
    rng = np.random.default_rng(8)
    lens = np.array([25, 40, 9], dtype=np.int64)
    ids = np.zeros((3, 40), dtype=np.int64)
    for b, L in enumerate(lens):
        ids[b, :L] = rng.integers(4, 50, size=L)
    sid = np.array([3, 77, 108])
    rows = np.array([[0.667, 1.0, 0.8], [0.0, 1.3, 0.0], [0.3, 0.75, 1.0]], dtype=np.float32)
    mixed = sess.infer(ids, lens, None, sid, seed=21, row_scales=rows)
    for b in range(3):
        uni = sess.infer(ids, lens, tuple(rows[b]), sid, seed=21)
        assert uni.frames[b] == mixed.frames[b]
        np.testing.assert_array_equal(mixed.utterance_pcm(b), uni.utterance_pcm(b))
    assert mixed.frames[1] != sess.infer(ids, lens, (0.0, 1.0, 0.0), sid).frames[1]   # length_scale 1.3 took effect
    with pytest.raises(ValueError):
        sess.infer(ids, lens, None, sid)                                                # neither scales nor row_scales
    with pytest.raises(ValueError, match="row_scales"):
        sess.infer(ids, lens, None, sid, row_scales=np.array([[0, 1, 0]] * 2, dtype=np.float32))
    sess.close()


@pytest.mark.gpu
def test_batched_queue_equals_sentence_by_sentence(built_library, voices):
    """all sentences of an utterance in one engine call == the reference's loop of one
    `ids_to_audio` per sentence followed by the host post chain (tts.py:519-551), result by result; and the
    device-assembled WAV == the HTTP server's host-side WAV assembly (mimic3_http/synthesis.py:60-85)."""
    from mimic3_b200 import tts
    from mimic3_b200.voice import B200Voice
    voice = B200Voice.load_from_directory(voices("tiny_ms"))
    settings = tts.B200Settings(voice="x/tiny", noise_scale=0.0, noise_w=0.0)

    def fill(q):
        q.speak_phonemes([["a", "b"], ["c"]])
        q.add_break(40)
        q.settings.speaker, q.settings.length_scale, q.settings.volume = "p201", 1.4, 35.0
        q.speak_phonemes([["d"], ["e", "f", "g"]])
        q.set_mark("m1")
        q.settings.rate, q.settings.volume = 1.25, 100.0
        q.speak_phonemes([["h", "a"]], is_utterance=False)
        q.speak_phonemes([["b"]])
        q.add_break(5)

    q = tts.B200UtteranceQueue(tts.B200Settings(**vars(settings)), lambda key: voice)
    fill(q)
    plan = tts.plan_sentences(q._results)
    end_settings = tts.B200Settings(**vars(q.settings))
    got = list(q.end_utterance())
    assert len(got) == len(plan)
    n_sent = 0
    for g, item in zip(got, plan):
        if isinstance(item, tts._Sentence):   # the reference's per-sentence path on the same engine
            n_sent += 1
            s = item.settings or end_settings
            audio = voice.ids_to_audio(voice.phonemes_to_ids(item.phonemes), speaker=s.speaker, length_scale=s.length_scale,
                                       noise_scale=s.noise_scale, noise_w=s.noise_w, rate=s.rate)
            want = audio if s.volume == 100.0 else pc.audioop_mul_int16(audio, s.volume / 100.0)
            assert isinstance(g, tts.AudioResult) and g.audio_bytes == want.tobytes()
        else:
            assert g is item
    assert n_sent == 3
    q2 = tts.B200UtteranceQueue(tts.B200Settings(**vars(settings)), lambda key: voice)
    fill(q2)
    assert q2.end_utterance_wav() == tts.results_to_wav_bytes(got)
