"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): identical integer durations, waveform RMS error <= 1e-3
per utterance against the fp32 reference restatement; int16 PCM within 1 LSB wherever
the float waveform matches to fp32 accuracy.
"""
import numpy as np
import pytest

from conftest import rand_ids

pytestmark = pytest.mark.gpu

RMS_TOL = 1e-3  # north_star: "waveform RMS error vs fp32 reference <= 1e-3"


@pytest.fixture(scope="module")
def sessions(built_library, voices):
    from mimic3_b200.engine import B200Session
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = B200Session(str(voices(name)))
        return cache[name]
    yield get
    for s in cache.values():
        s.close()


@pytest.fixture(scope="module")
def oracles(voices):
    from oracle.vits_oracle import VitsOracle
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = VitsOracle(str(voices(name)))
        return cache[name]
    return get


def _batch(rng, nsym, lengths):
    T = max(lengths)
    ids = np.zeros((len(lengths), T), dtype=np.int64)
    for b, L in enumerate(lengths):
        ids[b, :L] = rand_ids(rng, nsym, L)
    return ids, np.array(lengths, dtype=np.int64)


def _compare(sess, orc, ids, lengths, scales, sid=None, seed=0, tol=RMS_TOL):
    from oracle.vits_oracle import audio_float_to_int16
    r = sess.infer(ids, lengths, scales, sid, seed=seed, keep_float=True,
                   debug_tensors=("durations", "x", "stats", "logw", "z_p", "z"))
    worst = 0.0
    tok = 0
    for b, L in enumerate(lengths):
        audio, inter = orc.infer(ids[b, :L], scales, sid=None if sid is None else int(sid[b]), seed=seed, row=b,
                                 return_intermediates=True)
        dur = r.tensors["durations"][tok:tok + L, 0].astype(np.int64)
        np.testing.assert_array_equal(dur, inter["durations"], err_msg=f"durations differ (utt {b})")
        assert r.frames[b] == max(1, int(inter["durations"].sum()))
        got = r.utterance_audio(b)
        assert got.shape == audio.shape
        rms = float(np.sqrt(np.mean((got - audio) ** 2)))
        worst = max(worst, rms)
        assert rms <= tol, f"utt {b}: waveform RMS {rms:.3e} > {tol}"
        np.testing.assert_allclose(r.tensors["x"][tok:tok + L], inter["x"], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(r.tensors["logw"][tok:tok + L, 0], inter["logw"], atol=2e-4, rtol=1e-4)
        # projected prior statistics (m_p | logs_p), the expanded + noised prior z_p and the flow output z:
        # token rows / frame rows of this utterance in the packed, channels-last engine tensors
        np.testing.assert_allclose(r.tensors["stats"][tok:tok + L], np.concatenate([inter["m_p"], inter["logs_p"]], 1),
                                   atol=3e-4, rtol=1e-4, err_msg=f"stats (utt {b})")
        f0 = int(r.frames[:b].sum())
        zp, z = r.tensors["z_p"][f0:f0 + r.frames[b]], r.tensors["z"][f0:f0 + r.frames[b]]
        assert zp.shape == inter["z_p"].shape and z.shape == inter["z"].shape
        np.testing.assert_allclose(zp, inter["z_p"], atol=1e-3, rtol=1e-4, err_msg=f"z_p (utt {b})")
        rel = float(np.sqrt(np.mean((z - inter["z"]) ** 2)) / max(1e-12, np.sqrt(np.mean(inter["z"] ** 2))))
        assert rel <= 5e-3, f"utt {b}: flow output z relative RMS {rel:.3e}"  # fp16 operands, fp32 accumulate
        # the engine's own int16 conversion is bit-exact w.r.t. the reference formula on ITS float audio
        np.testing.assert_array_equal(r.utterance_pcm(b), audio_float_to_int16(got))
        assert abs(float(r.peaks[b]) - float(np.abs(got).max())) == 0.0
        tok += L
    return worst


@pytest.mark.parametrize("voice,multi", [("tiny", False), ("tiny_ms", True), ("tiny_rb1_dp", True)])
def test_tiny_voices_deterministic_parity(sessions, oracles, voice, multi):
    rng = np.random.default_rng(7)
    sess, orc = sessions(voice), oracles(voice)
    lengths = [17, 1, 40, 5, 9]
    ids, lens = _batch(rng, sess.info.num_symbols, lengths)
    sid = np.array([b % sess.info.n_speakers for b in range(len(lengths))]) if multi else None
    worst = _compare(sess, orc, ids, lens, (0.0, 1.0, 0.0), sid)
    print(f"{voice}: worst RMS {worst:.3e}")
    worst = _compare(sess, orc, ids, lens, (0.0, 1.3, 0.0), sid)


def test_export_styles_give_identical_audio(sessions):
    rng = np.random.default_rng(3)
    ids, lens = _batch(rng, 20, [12, 30])
    sid = np.array([1, 2])
    outs = [sessions(v).infer(ids, lens, (0.0, 1.0, 0.0), sid, keep_float=True) for v in
            ("tiny_ms", "tiny_ms_folded", "tiny_ms_wn")]
    np.testing.assert_array_equal(outs[0].audio, outs[1].audio)
    assert outs[0].audio.shape == outs[2].audio.shape
    # weight_g/weight_v reconstruction differs by an fp32 ulp, which can flip an fp16 operand rounding
    assert np.sqrt(np.mean((outs[0].audio - outs[2].audio) ** 2)) < 5e-4


def test_noise_on_parity_with_shared_philox(sessions, oracles):
    """Noise > 0: reference RNG is irreproducible; engine and oracle share the Philox spec."""
    rng = np.random.default_rng(9)
    sess, orc = sessions("tiny_ms"), oracles("tiny_ms")
    ids, lens = _batch(rng, 20, [21, 8, 33])
    _compare(sess, orc, ids, lens, (0.667, 1.0, 0.8), np.array([0, 1, 2]), seed=1234)
    a = sess.infer(ids, lens, (0.667, 1.0, 0.8), np.array([0, 1, 2]), seed=1).pcm
    b = sess.infer(ids, lens, (0.667, 1.0, 0.8), np.array([0, 1, 2]), seed=2).pcm
    assert a.shape != b.shape or not np.array_equal(a, b)


def test_low_voice_parity_single_and_multispeaker(sessions, oracles):
    rng = np.random.default_rng(11)
    for voice, multi in (("low", False), ("low_ms", True)):
        sess, orc = sessions(voice), oracles(voice)
        ids, lens = _batch(rng, sess.info.num_symbols, [60, 23])
        sid = np.array([0, 57]) if multi else None
        worst = _compare(sess, orc, ids, lens, (0.0, 1.0, 0.0), sid)
        print(f"{voice}: worst RMS {worst:.3e}")


def test_batch_rows_equal_single_runs(sessions):
    """Batch-1 edge semantics inside a ragged batch (SURVEY.md §7 hard part 3)."""
    rng = np.random.default_rng(5)
    sess = sessions("tiny_ms")
    lengths = [9, 31, 2, 17]
    ids, lens = _batch(rng, 20, lengths)
    sid = np.array([2, 0, 1, 1])
    r = sess.infer(ids, lens, (0.0, 1.0, 0.0), sid, keep_float=True)
    for b, L in enumerate(lengths):
        one = sess.infer(ids[b:b + 1, :L], lens[b:b + 1], (0.0, 1.0, 0.0), sid[b:b + 1], keep_float=True)
        np.testing.assert_array_equal(one.audio, r.utterance_audio(b))
        np.testing.assert_array_equal(one.pcm, r.utterance_pcm(b))


def test_ort_style_run_and_errors(sessions):
    sess = sessions("tiny_ms")
    feed = {"input": np.array([[4, 5, 6, 7]], dtype=np.int64), "input_lengths": np.array([4], dtype=np.int64),
            "scales": np.array([0.0, 1.0, 0.0], dtype=np.float32), "sid": np.array([1], dtype=np.int64)}
    out = sess.run(None, feed)
    assert len(out) == 1 and out[0].dtype == np.float32 and out[0].ndim == 3 and out[0].shape[:2] == (1, 1)
    audio = out[0].squeeze()                       # voice.py:230
    assert audio.size % sess.info.hop_length == 0 and np.abs(audio).max() <= 1.0
    bad = dict(feed, input=np.array([[4, 99, 6, 7]], dtype=np.int64))
    with pytest.raises(ValueError, match="outside"):
        sess.run(None, bad)
    with pytest.raises(ValueError, match="sid"):
        sess.run(None, dict(feed, sid=np.array([3], dtype=np.int64)))
    with pytest.raises(ValueError):
        sess.run(None, {k: v for k, v in feed.items() if k != "sid"})
    with pytest.raises(ValueError):
        sess.run(None, dict(feed, input_lengths=np.array([0], dtype=np.int64)))
    with pytest.raises(ValueError):
        sess.run(None, dict(feed, input_lengths=np.array([5], dtype=np.int64)))
    ok = sess.run(None, feed)                     # the engine survives bad calls
    np.testing.assert_array_equal(ok[0], out[0])


def test_thread_safety_shared_voice(sessions):
    """One voice shared by worker threads, like the shared ORT session (voice.py:277-292)."""
    import threading
    sess = sessions("tiny_ms")
    rng = np.random.default_rng(2)
    jobs = [_batch(rng, 20, [int(rng.integers(3, 40)) for _ in range(3)]) for _ in range(8)]
    want = [sess.infer(i, l, (0.0, 1.0, 0.0), np.zeros(3, dtype=np.int64)).pcm for i, l in jobs]
    got = [None] * len(jobs)

    def work(k):
        for _ in range(3):
            got[k] = sess.infer(jobs[k][0], jobs[k][1], (0.0, 1.0, 0.0), np.zeros(3, dtype=np.int64)).pcm
    ts = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for w, g in zip(want, got):
        np.testing.assert_array_equal(w, g)


def test_full_size_properties_cfg2(sessions):
    """BASELINE cfg2 shape (vctk_low-like, B=32, 100 ids, sid 0): size-independent properties."""
    sess = sessions("low_ms")
    rng = np.random.Generator(np.random.PCG64(1234))
    ids = rng.integers(4, sess.info.num_symbols, size=(32, 100)).astype(np.int64)
    lens = np.full(32, 100, dtype=np.int64)
    sid = np.zeros(32, dtype=np.int64)
    r = sess.infer(ids, lens, (0.0, 1.0, 0.0), sid, keep_float=True)
    hop = sess.info.hop_length
    assert np.all(np.diff(r.sample_offsets) == r.frames * hop)
    for b in range(32):
        pcm = r.utterance_pcm(b)
        # peak*(32767/peak) in fp32 lands on 32767 or one ulp below it (-> 32766 after truncation)
        assert int(np.abs(pcm.astype(np.int32)).max()) in (32766, 32767)
        assert np.isfinite(r.utterance_audio(b)).all() and np.abs(r.utterance_audio(b)).max() <= 1.0
    r2 = sess.infer(ids, lens, (0.0, 1.0, 0.0), sid)
    np.testing.assert_array_equal(r.pcm, r2.pcm)                       # idempotent / deterministic
    sub = sess.infer(ids[5:7], lens[5:7], (0.0, 1.0, 0.0), sid[5:7])
    np.testing.assert_array_equal(sub.pcm, r.pcm[r.sample_offsets[5]:r.sample_offsets[7]])
    half = sess.infer(ids, lens, (0.0, 0.5, 0.0), sid)
    assert half.total_samples < r.total_samples                        # length_scale monotone
    assert r.launches > 0


def test_b200voice_end_to_end(voices, built_library):
    from mimic3_b200.voice import B200Voice
    v = B200Voice.load_from_directory(voices("tiny_ms"))
    v2 = B200Voice.load_from_directory(voices("tiny_ms"))
    assert v.onnx_model is v2.onnx_model                                # shared per generator.onnx
    a = v.ids_to_audio([4, 5, 6, 7, 8], speaker="p201", noise_scale=0.0, noise_w=0.0)
    b = v.ids_to_audio_batch([[4, 5, 6, 7, 8], [9, 10]], speakers=["p201", "p200"], noise_scale=0.0, noise_w=0.0)
    assert a.dtype == np.int16 and np.array_equal(a, b[0])
    assert int(np.abs(a.astype(np.int32)).max()) in (32766, 32767)


def test_tensor_core_mrf_matches_oracle_and_simt(voices, built_library, oracles, monkeypatch):
    """tcgen05 MRF (bf16 and fp16 operands, fp32 TMEM residual) vs the fp32 oracle, stage by stage,
    and vs the fp32 SIMT path of the same library."""
    from mimic3_b200.engine import B200Session
    rng = np.random.default_rng(21)
    orc = oracles("low_ms")
    ids, lens = _batch(rng, 50, [37, 70, 3])
    sid = np.array([5, 100, 42])
    monkeypatch.setenv("M3B200_FORCE_SIMT", "1")
    simt = B200Session(str(voices("low_ms")))
    monkeypatch.delenv("M3B200_FORCE_SIMT")
    names = ("z", "mrf0", "mrf1", "mrf2")
    ref = simt.infer(ids, lens, (0.0, 1.0, 0.0), sid, keep_float=True, debug_tensors=names)
    for fmt in ("bf16", "fp16"):
        monkeypatch.setenv("M3B200_TC_FORMAT", fmt)
        sess = B200Session(str(voices("low_ms")))
        r = sess.infer(ids, lens, (0.0, 1.0, 0.0), sid, keep_float=True, debug_tensors=names)
        np.testing.assert_array_equal(r.frames, ref.frames)
        for name in names:
            if name not in r.tensors:   # the fused last stage never materialises its MRF output
                assert name == "mrf2"
                continue
            a, b = r.tensors[name], ref.tensors[name]
            rel = np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))
            print(f"{fmt} {name}: relative RMS vs SIMT fp32 {rel:.3e}")
            assert rel < (5e-2 if fmt == "bf16" else 8e-3), (fmt, name, rel)
        off = 0
        for b, L in enumerate(lens):
            audio, inter = orc.infer(ids[b, :L], (0.0, 1.0, 0.0), sid=int(sid[b]), return_intermediates=True)
            got = r.utterance_audio(b)
            rms = float(np.sqrt(np.mean((got - audio) ** 2)))
            print(f"{fmt} utt {b}: waveform RMS vs oracle {rms:.3e} (signal RMS {np.sqrt(np.mean(audio**2)):.3f})")
            assert rms <= (RMS_TOL if fmt == "fp16" else 5 * RMS_TOL), (fmt, b, rms)  # fp16 is the shipped default
        sess.close()
    simt.close()


def test_long_form_cfg4_shape(sessions, oracles):
    """BASELINE configs[3]: long-form utterances (~1800 ids), length_scale in {0.8, 1.0, 1.2}.
    Parity vs the oracle at 600 ids (seconds on CPU); properties at the full 1800."""
    sess, orc = sessions("low"), oracles("low")
    rng = np.random.default_rng(31)
    ids, lens = _batch(rng, sess.info.num_symbols, [600])
    worst = _compare(sess, orc, ids, lens, (0.0, 1.2, 0.0), None)
    print(f"long-form 600 ids: worst RMS {worst:.3e}")
    ids, lens = _batch(rng, sess.info.num_symbols, [1800, 1750, 900, 1800])
    frames = {}
    for ls in (0.8, 1.0, 1.2):
        r = sess.infer(ids, lens, (0.0, ls, 0.0), None, keep_float=True)
        frames[ls] = r.frames.copy()
        assert np.isfinite(r.audio).all() and np.abs(r.audio).max() <= 1.0
        assert np.all(np.diff(r.sample_offsets) == r.frames * sess.info.hop_length)
        one = sess.infer(ids[2:3, :900], lens[2:3], (0.0, ls, 0.0), None)       # row 2 alone == row 2 in the batch
        np.testing.assert_array_equal(one.pcm, r.utterance_pcm(2))
    assert np.all(frames[0.8] <= frames[1.0]) and np.all(frames[1.0] <= frames[1.2])


def test_mixed_voice_residency_cfg5(voices, built_library):
    """BASELINE configs[4]: several voices resident at once, rows grouped by voice; switching voice is
    a pointer swap and must not disturb results."""
    from mimic3_b200.engine import B200Session
    names = ["low_ms", "low", "tiny_ms"]
    sess = {n: B200Session(str(voices(n))) for n in names}
    rng = np.random.default_rng(41)
    jobs = {}
    for n in names:
        ns = sess[n].info.num_symbols
        ids, lens = _batch(rng, ns, [int(rng.integers(10, 60)) for _ in range(5)])
        sid = (np.arange(5) % sess[n].info.n_speakers) if sess[n].info.has_speaker_embedding else None
        jobs[n] = (ids, lens, sid)
    first = {n: sess[n].infer(*jobs[n][:2], (0.0, 1.0, 0.0), jobs[n][2]).pcm for n in names}
    for _ in range(2):  # interleave voices
        for n in reversed(names):
            again = sess[n].infer(*jobs[n][:2], (0.0, 1.0, 0.0), jobs[n][2]).pcm
            np.testing.assert_array_equal(again, first[n])
    for s in sess.values():
        s.close()


def test_text_side_tensor_core_split_keeps_durations(voices, built_library, monkeypatch):
    """The fp16 hi/lo split GEMMs (3 MMAs per product) of the text encoder / duration predictor must
    reproduce the fp32 FFMA path: identical integer durations on 256 x 80 ids, logw within fp32 noise."""
    from mimic3_b200.engine import B200Session
    rng = np.random.Generator(np.random.PCG64(1234))
    ids = rng.integers(4, 50, size=(256, 80)).astype(np.int64)
    lens = np.full(256, 80, dtype=np.int64)
    sid = (np.arange(256) % 109).astype(np.int64)
    monkeypatch.setenv("M3B200_TEXT_SIMT", "1")
    ref_s = B200Session(str(voices("low_ms")))
    monkeypatch.delenv("M3B200_TEXT_SIMT")
    tc_s = B200Session(str(voices("low_ms")))
    for scales, seed in (((0.0, 1.0, 0.0), 0), ((0.667, 1.0, 0.8), 99)):
        a = ref_s.infer(ids, lens, scales, sid, seed=seed, debug_tensors=("durations", "logw", "x"))
        b = tc_s.infer(ids, lens, scales, sid, seed=seed, debug_tensors=("durations", "logw", "x"))
        dl = np.abs(a.tensors["logw"] - b.tensors["logw"]).max()
        dx = np.abs(a.tensors["x"] - b.tensors["x"]).max()
        flips = int((a.tensors["durations"] != b.tensors["durations"]).sum())
        print(f"scales {scales}: max|dlogw| {dl:.2e}  max|dx| {dx:.2e}  duration flips {flips} / {a.tensors['durations'].size}")
        assert flips == 0
        assert dl < 5e-5 and dx < 5e-4
        np.testing.assert_array_equal(a.frames, b.frames)
    ref_s.close()
    tc_s.close()


def test_persistent_last_stage_matches_per_window_kernel_and_oracle(voices, built_library, oracles, monkeypatch):
    """The persistent warp-specialised last-stage kernels (phase-major planes, kernels_tc_dec3.cu = default; sample-order
    windows, kernels_tc_dec2.cu) against the per-window fused kernel, the unfused kernels and the oracle, on a ragged
    batch that mixes one-window utterances, many-window utterances and more work items than SMs."""
    from mimic3_b200.engine import B200Session
    rng = np.random.default_rng(77)
    orc = oracles("low_ms")
    lens_list = [1, 2, 5, 80, 33, 64] + [int(x) for x in rng.integers(3, 90, size=26)]
    ids, lens = _batch(rng, 50, lens_list)
    sid = rng.integers(0, 109, size=len(lens_list))
    scales = (0.0, 1.0, 0.0)
    sess = B200Session(str(voices("low_ms")))
    got = sess.infer(ids, lens, scales, sid, keep_float=True)
    again = sess.infer(ids, lens, scales, sid, keep_float=True)
    monkeypatch.setenv("M3B200_DEC_V2", "1")
    v2 = sess.infer(ids, lens, scales, sid, keep_float=True)
    monkeypatch.delenv("M3B200_DEC_V2")
    monkeypatch.setenv("M3B200_DEC_V1", "1")
    v1 = sess.infer(ids, lens, scales, sid, keep_float=True)
    monkeypatch.delenv("M3B200_DEC_V1")
    monkeypatch.setenv("M3B200_UNFUSED_DEC", "1")
    unf = sess.infer(ids, lens, scales, sid, keep_float=True)
    monkeypatch.delenv("M3B200_UNFUSED_DEC")
    monkeypatch.setenv("M3B200_MRF_V1", "1")   # per-window MRF kernel for the C = 64 stage instead of the persistent one
    mrf_v1 = sess.infer(ids, lens, scales, sid, keep_float=True)
    monkeypatch.delenv("M3B200_MRF_V1")
    np.testing.assert_array_equal(got.frames, v1.frames)
    np.testing.assert_array_equal(got.audio, again.audio)   # deterministic: no order-dependent accumulation
    np.testing.assert_array_equal(got.peaks, again.peaks)
    for other, name in ((v2, "sample-order persistent"), (v1, "per-window fused"), (unf, "unfused"), (mrf_v1, "per-window MRF")):
        for b in range(len(lens_list)):
            a, c = got.utterance_audio(b), other.utterance_audio(b)
            rms = float(np.sqrt(np.mean((a - c) ** 2)))
            assert rms < 3e-4, (name, b, rms)
            assert abs(float(got.peaks[b]) - float(np.abs(a).max())) == 0.0
    for b in (0, 1, 2, 3, 10):
        audio = orc.infer(ids[b, :lens[b]], scales, sid=int(sid[b]))
        rms = float(np.sqrt(np.mean((got.utterance_audio(b) - audio) ** 2)))
        print(f"utt {b} ({lens[b]} ids): RMS vs oracle {rms:.3e}")
        assert rms <= RMS_TOL
    sess.close()


def test_short_attention_matches_generic_kernel(voices, built_library, monkeypatch):
    """attention_short_kernel (whole utterance/head in shared memory, kernels_attn.cu) against the generic
    flash-style kernel on ragged batches: one query block per utterance (max 128 ids), several query blocks
    (max 200 ids), and the tiny voice (dk = 16).  Same encoder output to fp32 noise, identical durations."""
    from mimic3_b200.engine import B200Session
    rng = np.random.default_rng(91)
    cases = [("low_ms", [1, 2, 5, 9, 31, 64, 80, 100, 128, 77]), ("low_ms", [200, 3, 150, 96]),
             ("tiny_ms", [40, 1, 17, 33, 8])]
    for voice, lens_list in cases:
        sess = B200Session(str(voices(voice)))
        ids, lens = _batch(rng, sess.info.num_symbols, lens_list)
        sid = rng.integers(0, sess.info.n_speakers, size=len(lens_list))
        names = ("x", "logw", "durations")
        new = sess.infer(ids, lens, (0.0, 1.0, 0.0), sid, debug_tensors=names)
        monkeypatch.setenv("M3B200_ATTN_V1", "1")
        old = sess.infer(ids, lens, (0.0, 1.0, 0.0), sid, debug_tensors=names)
        monkeypatch.delenv("M3B200_ATTN_V1")
        dx = float(np.abs(new.tensors["x"] - old.tensors["x"]).max())
        dl = float(np.abs(new.tensors["logw"] - old.tensors["logw"]).max())
        print(f"{voice} {lens_list}: max|dx| {dx:.2e}  max|dlogw| {dl:.2e}  launches {new.launches} vs {old.launches}")
        assert dx < 5e-5 and dl < 5e-5
        np.testing.assert_array_equal(new.tensors["durations"], old.tensors["durations"])
        np.testing.assert_array_equal(new.frames, old.frames)
        for nb in ("1", "3"):   # forced query-block counts exercise the multi-block path on short rows too
            monkeypatch.setenv("M3B200_ATTN_NB", nb)
            alt = sess.infer(ids, lens, (0.0, 1.0, 0.0), sid, debug_tensors=names)
            monkeypatch.delenv("M3B200_ATTN_NB")
            assert float(np.abs(alt.tensors["x"] - old.tensors["x"]).max()) < 5e-5
        sess.close()


def test_rowgemm_variants_keep_text_side_results(voices, built_library, monkeypatch):
    """128-column CTAs for the 1x1 layers and the 3-stage ring of rowgemm_tc_kernel are pure re-tilings:
    same products in the same order per output, so the text side must not move at all."""
    from mimic3_b200.engine import B200Session
    rng = np.random.Generator(np.random.PCG64(5))
    ids = rng.integers(4, 50, size=(24, 80)).astype(np.int64)
    lens = np.full(24, 80, dtype=np.int64)
    lens[3], lens[7] = 11, 1
    sid = (np.arange(24) % 109).astype(np.int64)
    names = ("x", "stats", "logw", "durations")
    base = B200Session(str(voices("low_ms")))
    ref = base.infer(ids, lens, (0.667, 1.0, 0.8), sid, seed=5, debug_tensors=names)
    for env in ({"M3B200_ROWGEMM_NC": "128"}, {"M3B200_ROWGEMM_STAGES": "3"},
                {"M3B200_ROWGEMM_NC": "128", "M3B200_ROWGEMM_STAGES": "3"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        alt_s = B200Session(str(voices("low_ms")))
        alt = alt_s.infer(ids, lens, (0.667, 1.0, 0.8), sid, seed=5, debug_tensors=names)
        for k in env:
            monkeypatch.delenv(k)
        for n in names:
            np.testing.assert_array_equal(alt.tensors[n], ref.tensors[n], err_msg=f"{env}: {n}")
        np.testing.assert_array_equal(alt.pcm, ref.pcm)
        alt_s.close()
    base.close()


def test_wide_io_is_bit_identical(sessions, monkeypatch):
    """256-bit global loads/stores in the row GEMM and the polyphase upsampler epilogue (default) against the 128-bit
    variant (M3B200_WIDE_IO=0): only the width of memory instructions changes, every output bit must stay the same."""
    rng = np.random.default_rng(123)
    sess = sessions("low_ms")
    ids, lens = _batch(rng, 50, [80, 3, 41, 1, 64, 17])
    sid = np.array([0, 5, 108, 7, 33, 2])
    for scales, seed in (((0.0, 1.0, 0.0), 0), ((0.667, 1.1, 0.8), 4)):
        ref = sess.infer(ids, lens, scales, sid, seed=seed, keep_float=True, debug_tensors=("x", "logw", "z"))
        monkeypatch.setenv("M3B200_WIDE_IO", "0")
        alt = sess.infer(ids, lens, scales, sid, seed=seed, keep_float=True, debug_tensors=("x", "logw", "z"))
        monkeypatch.delenv("M3B200_WIDE_IO")
        for n in ("x", "logw", "z"):
            np.testing.assert_array_equal(alt.tensors[n], ref.tensors[n], err_msg=n)
        np.testing.assert_array_equal(alt.frames, ref.frames)
        np.testing.assert_array_equal(alt.audio, ref.audio)
        np.testing.assert_array_equal(alt.pcm, ref.pcm)
    tiny = sessions("tiny_ms")
    ids, lens = _batch(rng, 20, [9, 30])
    ref = tiny.infer(ids, lens, (0.0, 1.0, 0.0), np.array([0, 1]), keep_float=True)
    monkeypatch.setenv("M3B200_WIDE_IO", "0")
    alt = tiny.infer(ids, lens, (0.0, 1.0, 0.0), np.array([0, 1]), keep_float=True)
    monkeypatch.delenv("M3B200_WIDE_IO")
    np.testing.assert_array_equal(alt.audio, ref.audio)


def test_upsampler_kernel_is_bit_identical_to_generic_conv(sessions, monkeypatch):
    """ups_tc_kernel (kernels_tc_ups.cu: bias in shared memory, early accumulator release; the default since round 2)
    computes exactly what conv_tc_kernel's TC_UPS epilogue (M3B200_UPS_V1=1) computes: same MMAs, same `acc + bias`."""
    rng = np.random.default_rng(2)
    for voice, nsym, lens_list in (("low_ms", 50, [80, 3, 41, 1, 64, 17, 100]), ("tiny_ms", 20, [9, 30, 1])):
        sess = sessions(voice)
        ids, lens = _batch(rng, nsym, lens_list)
        sid = rng.integers(0, sess.info.n_speakers, size=len(lens_list))
        new = sess.infer(ids, lens, (0.667, 1.0, 0.8), sid, seed=3, keep_float=True)
        monkeypatch.setenv("M3B200_UPS_V1", "1")
        old = sess.infer(ids, lens, (0.667, 1.0, 0.8), sid, seed=3, keep_float=True)
        monkeypatch.delenv("M3B200_UPS_V1")
        np.testing.assert_array_equal(new.frames, old.frames)
        np.testing.assert_array_equal(new.audio, old.audio)
        np.testing.assert_array_equal(new.pcm, old.pcm)


def _compare_rows(sess, orc, ids, lens, sid, scales, seed, rows, tag):
    """Whole batch through the engine once, the chosen rows one by one through the oracle (batch 1, same Philox
    stream selected by `row`).  Durations equal, waveform RMS <= 1e-3, PCM == the reference formula on the engine's
    float audio and within 1 LSB of the oracle's PCM wherever the float waveforms agree to 1e-6."""
    from oracle.vits_oracle import audio_float_to_int16
    r = sess.infer(ids, lens, scales, sid, seed=seed, keep_float=True, debug_tensors=("durations", "logw", "z_p", "z", "stats"))
    tok_off = np.concatenate([[0], np.cumsum(lens)])
    frm_off = np.concatenate([[0], np.cumsum(r.frames)])
    worst = 0.0
    for b in rows:
        L = int(lens[b])
        audio, inter = orc.infer(ids[b, :L], scales, sid=None if sid is None else int(sid[b]), seed=seed, row=b,
                                 return_intermediates=True)
        t0, f0 = int(tok_off[b]), int(frm_off[b])
        np.testing.assert_array_equal(r.tensors["durations"][t0:t0 + L, 0].astype(np.int64), inter["durations"],
                                      err_msg=f"{tag}: durations differ (row {b})")
        assert r.frames[b] == max(1, int(inter["durations"].sum()))
        np.testing.assert_allclose(r.tensors["logw"][t0:t0 + L, 0], inter["logw"], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(r.tensors["stats"][t0:t0 + L], np.concatenate([inter["m_p"], inter["logs_p"]], 1),
                                   atol=3e-4, rtol=1e-4)
        np.testing.assert_allclose(r.tensors["z_p"][f0:f0 + r.frames[b]], inter["z_p"], atol=1e-3, rtol=1e-4)
        z = r.tensors["z"][f0:f0 + r.frames[b]]
        rel = float(np.sqrt(np.mean((z - inter["z"]) ** 2)) / np.sqrt(np.mean(inter["z"] ** 2)))
        assert rel <= 5e-3, f"{tag} row {b}: z relative RMS {rel:.3e}"
        got = r.utterance_audio(b)
        assert got.shape == audio.shape
        rms = float(np.sqrt(np.mean((got - audio) ** 2)))
        worst = max(worst, rms)
        assert rms <= RMS_TOL, f"{tag} row {b}: waveform RMS {rms:.3e}"
        pcm = r.utterance_pcm(b)
        np.testing.assert_array_equal(pcm, audio_float_to_int16(got))
        want = audio_float_to_int16(audio)
        close = np.abs(got - audio) <= 1e-6
        if abs(float(np.abs(got).max()) - float(np.abs(audio).max())) <= 1e-6 and close.any():
            assert np.abs(pcm[close].astype(np.int32) - want[close].astype(np.int32)).max() <= 1
    return worst, r


def test_benchmarked_config3_parity_256x80_all_speakers(sessions, oracles):
    """The exact batch bench.py times (BASELINE configs[2]: `bench.make_inputs()`: 256 x 80 ids, PCG64(1234),
    sid = b mod 109, vctk_low-shaped voice) against the oracle: 20 rows spread over the batch and the speakers, at
    the parity settings (0, 1, 0) and at the voice defaults (0.667, 1, 0.8) with the shared Philox streams."""
    import bench
    sess, orc = sessions("low_ms"), oracles("low_ms")
    ids, lens, sid = bench.make_inputs()
    assert ids.shape == (256, 80) and int(sid.max()) == 108
    rows = sorted(set(int(x) for x in np.linspace(0, 255, 18)) | {108, 109})
    for scales, seed in (((0.0, 1.0, 0.0), 0), ((0.667, 1.0, 0.8), 2001)):
        worst, r = _compare_rows(sess, orc, ids, lens, sid, scales, seed, rows, f"cfg3 {scales}")
        print(f"cfg3 {scales}: {len(rows)} rows, worst RMS {worst:.3e}, frames/id {r.frames.sum() / lens.sum():.2f}")


def test_benchmarked_config2_parity_32x100_speaker0(sessions, oracles):
    """BASELINE configs[1]: batch 32 x 100 ids, speaker p239 = index 0 (voices.json:735-736): 16 rows vs the oracle,
    noise off and on."""
    sess, orc = sessions("low_ms"), oracles("low_ms")
    rng = np.random.Generator(np.random.PCG64(1234))
    ids = rng.integers(4, 50, size=(32, 100)).astype(np.int64)
    lens = np.full(32, 100, dtype=np.int64)
    sid = np.zeros(32, dtype=np.int64)
    rows = list(range(0, 32, 2))
    for scales, seed in (((0.0, 1.0, 0.0), 0), ((0.667, 1.0, 0.8), 77)):
        worst, r = _compare_rows(sess, orc, ids, lens, sid, scales, seed, rows, f"cfg2 {scales}")
        print(f"cfg2 {scales}: {len(rows)} rows, worst RMS {worst:.3e}")


def test_flow_second_generation_kernel_matches_first_and_unfused(sessions, oracles, monkeypatch):
    """flow2_tc_kernel (kernels_tc_flow2.cu: post(skip) folded into pre-multiplied skip weights, N = 96 / 192 MMAs) against
    the first fused kernel (M3B200_FLOW_V1=1), the per-conv kernels (M3B200_UNFUSED_FLOW=1) and the oracle, on a ragged
    batch with one-window, many-window and one-frame-tail utterances, noise on (per-utterance conditioning in play)."""
    rng = np.random.default_rng(2025)
    sess, orc = sessions("low_ms"), oracles("low_ms")
    lens_list = [1, 2, 27, 28, 80, 55, 3, 64] + [int(x) for x in rng.integers(4, 90, size=12)]
    ids, lens = _batch(rng, 50, lens_list)
    sid = rng.integers(0, 109, size=len(lens_list))
    for scales, seed in (((0.0, 1.0, 0.0), 0), ((0.667, 1.0, 0.8), 7)):
        new = sess.infer(ids, lens, scales, sid, seed=seed, keep_float=True, debug_tensors=("z_p", "z"))
        monkeypatch.setenv("M3B200_FLOW_V1", "1")
        v1 = sess.infer(ids, lens, scales, sid, seed=seed, keep_float=True, debug_tensors=("z_p", "z"))
        monkeypatch.delenv("M3B200_FLOW_V1")
        monkeypatch.setenv("M3B200_UNFUSED_FLOW", "1")
        unf = sess.infer(ids, lens, scales, sid, seed=seed, keep_float=True, debug_tensors=("z_p", "z"))
        monkeypatch.delenv("M3B200_UNFUSED_FLOW")
        np.testing.assert_array_equal(new.frames, v1.frames)
        np.testing.assert_array_equal(new.tensors["z_p"], v1.tensors["z_p"])
        zn, z1, zu = new.tensors["z"], v1.tensors["z"], unf.tensors["z"]
        ref = float(np.sqrt(np.mean(zu ** 2)))
        for other, name in ((z1, "first fused kernel"), (zu, "unfused")):
            rel = float(np.sqrt(np.mean((zn - other) ** 2))) / ref
            print(f"scales {scales}: flow2 vs {name}: relative RMS {rel:.2e}")
            assert rel < 3e-3, (name, rel)
        for b in range(len(lens_list)):
            a, c = new.utterance_audio(b), v1.utterance_audio(b)
            # two independent 16-bit operand roundings of the flow (each within 5e-4 of the fp32 oracle at the waveform)
            assert float(np.sqrt(np.mean((a - c) ** 2))) < 6e-4, b
    for b in (0, 2, 4, 9):
        audio = orc.infer(ids[b, :lens[b]], (0.0, 1.0, 0.0), sid=int(sid[b]))
        got = sess.infer(ids[b:b + 1, :lens[b]], lens[b:b + 1], (0.0, 1.0, 0.0), sid[b:b + 1], keep_float=True).utterance_audio(0)
        rms = float(np.sqrt(np.mean((got - audio) ** 2)))
        assert rms <= RMS_TOL, (b, rms)
