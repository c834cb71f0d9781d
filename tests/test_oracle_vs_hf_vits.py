"""Independent anchor for the oracle: Hugging Face transformers' ``VitsModel``.

The arithmetic of the reference's hot path lives in onnxruntime + generator.onnx, neither of which exists
offline (DESIGN.md "parity unpinned").  transformers ships a complete, separately written implementation
of the same published VITS inference graph (``transformers/models/vits/modeling_vits.py``, a port of
jaywalnut310/vits used for MMS-TTS).  Here the SAME synthetic weights are loaded into both the oracle
(through the voice directory / generator.onnx, like the engine) and ``VitsModel`` (by parameter-name
mapping), both run the deterministic settings of the reference's golden samples (noise scales 0), and
durations + waveforms must agree to fp32 noise.  HF implements ResBlock1 generators only
(``x += conv2(lrelu(conv1(lrelu(x))))``); the shipped ``*_low`` voices use ResBlock2 (``x += conv(lrelu(x))``).
ResBlock2 is anchored on HF's code too: HF's ``HifiGanResidualBlock.forward`` runs UNMODIFIED with every
``convs2[i]`` module swapped for the exact inverse of the leaky ReLU that precedes it, which removes the second
conv and nothing else (``test_oracle_matches_transformers_vits_resblock2*``).
"""
import numpy as np
import pytest

from mimic3_b200 import synth_voice as sv


def _hf_model(cfg, params):
    import torch
    from transformers import VitsConfig, VitsModel

    multi = cfg.n_speakers > 1
    hcfg = VitsConfig(
        vocab_size=cfg.num_symbols, hidden_size=cfg.hidden_channels, num_hidden_layers=cfg.n_layers,
        num_attention_heads=cfg.n_heads, window_size=sv.ATTN_WINDOW, ffn_dim=cfg.filter_channels,
        ffn_kernel_size=cfg.kernel_size, flow_size=cfg.inter_channels, spectrogram_bins=17,
        use_stochastic_duration_prediction=cfg.use_sdp, num_speakers=cfg.n_speakers,
        speaker_embedding_size=cfg.gin_channels if multi else 0,
        upsample_initial_channel=cfg.upsample_initial_channel, upsample_rates=list(cfg.upsample_rates),
        upsample_kernel_sizes=list(cfg.upsample_kernel_sizes), resblock_kernel_sizes=list(cfg.resblock_kernel_sizes),
        resblock_dilation_sizes=[list(d) for d in cfg.resblock_dilation_sizes], leaky_relu_slope=0.1,
        depth_separable_channels=2, depth_separable_num_layers=sv.SDP_DDS_LAYERS,
        duration_predictor_flow_bins=sv.SDP_BINS, duration_predictor_tail_bound=5.0,
        duration_predictor_kernel_size=3, duration_predictor_num_flows=sv.SDP_N_FLOWS,
        duration_predictor_filter_channels=sv.DP_FILTER, prior_encoder_num_flows=sv.FLOW_N_FLOWS,
        prior_encoder_num_wavenet_layers=sv.FLOW_WN_LAYERS, wavenet_kernel_size=sv.FLOW_KERNEL, wavenet_dilation_rate=1,
        hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, layerdrop=0.0, layer_norm_eps=1e-5,
        sampling_rate=22050)
    model = VitsModel(hcfg).eval()
    rb2 = str(cfg.resblock) == "2"
    if rb2:
        # ResBlock2 out of HF's ResBlock1 code: conv2(lrelu(t)) becomes t when conv2 := lrelu^-1
        class InverseLeakyReLU(torch.nn.Module):
            def forward(self, t):
                return torch.where(t >= 0, t, t / 0.1)
        for blk in model.decoder.resblocks:
            for i in range(len(blk.convs2)):
                blk.convs2[i] = InverseLeakyReLU()
    sd = model.state_dict()
    used = set()

    def put(hf_name, value):
        assert hf_name in sd, hf_name
        v = torch.from_numpy(np.ascontiguousarray(value))
        assert tuple(sd[hf_name].shape) == tuple(v.shape), (hf_name, tuple(sd[hf_name].shape), tuple(v.shape))
        sd[hf_name] = v
        used.add(hf_name)

    def conv(hf, mine):
        put(hf + ".weight", params[mine + ".weight"])
        if mine + ".bias" in params:
            put(hf + ".bias", params[mine + ".bias"])

    def wn_conv(hf, mine):  # weight_norm parametrisation: g * v / ||v|| with g = ||v|| reproduces v
        w = params[mine + ".weight"]
        put(hf + ".parametrizations.weight.original0", np.sqrt((w ** 2).sum(axis=(1, 2), keepdims=True)))
        put(hf + ".parametrizations.weight.original1", w)
        put(hf + ".bias", params[mine + ".bias"])

    def norm(hf, mine):
        put(hf + ".weight", params[mine + ".gamma"])
        put(hf + ".bias", params[mine + ".beta"])

    def dds(hf, mine):
        for i in range(sv.SDP_DDS_LAYERS):
            conv(f"{hf}.convs_dilated.{i}", f"{mine}.convs_sep.{i}")
            conv(f"{hf}.convs_pointwise.{i}", f"{mine}.convs_1x1.{i}")
            norm(f"{hf}.norms_1.{i}", f"{mine}.norms_1.{i}")
            norm(f"{hf}.norms_2.{i}", f"{mine}.norms_2.{i}")

    put("text_encoder.embed_tokens.weight", params["enc_p.emb.weight"])
    for l in range(cfg.n_layers):
        a, h = f"enc_p.encoder.attn_layers.{l}", f"text_encoder.encoder.layers.{l}"
        for mine, hf in (("conv_q", "q_proj"), ("conv_k", "k_proj"), ("conv_v", "v_proj"), ("conv_o", "out_proj")):
            put(f"{h}.attention.{hf}.weight", params[f"{a}.{mine}.weight"][:, :, 0])   # 1x1 conv == Linear
            put(f"{h}.attention.{hf}.bias", params[f"{a}.{mine}.bias"])
        put(f"{h}.attention.emb_rel_k", params[f"{a}.emb_rel_k"])
        put(f"{h}.attention.emb_rel_v", params[f"{a}.emb_rel_v"])
        norm(f"{h}.layer_norm", f"enc_p.encoder.norm_layers_1.{l}")
        norm(f"{h}.final_layer_norm", f"enc_p.encoder.norm_layers_2.{l}")
        conv(f"{h}.feed_forward.conv_1", f"enc_p.encoder.ffn_layers.{l}.conv_1")
        conv(f"{h}.feed_forward.conv_2", f"enc_p.encoder.ffn_layers.{l}.conv_2")
    conv("text_encoder.project", "enc_p.proj")
    if cfg.use_sdp:
        conv("duration_predictor.conv_pre", "dp.pre")
        conv("duration_predictor.conv_proj", "dp.proj")
        dds("duration_predictor.conv_dds", "dp.convs")
        put("duration_predictor.flows.0.translate", params["dp.flows.0.m"])
        put("duration_predictor.flows.0.log_scale", params["dp.flows.0.logs"])
        for hf_i, mine_i in ((2, 3), (3, 5), (4, 7)):  # HF drops the Flip modules; its flows.1 is the unused vflow
            conv(f"duration_predictor.flows.{hf_i}.conv_pre", f"dp.flows.{mine_i}.pre")
            dds(f"duration_predictor.flows.{hf_i}.conv_dds", f"dp.flows.{mine_i}.convs")
            conv(f"duration_predictor.flows.{hf_i}.conv_proj", f"dp.flows.{mine_i}.proj")
    else:
        conv("duration_predictor.conv_1", "dp.conv_1")
        conv("duration_predictor.conv_2", "dp.conv_2")
        norm("duration_predictor.norm_1", "dp.norm_1")
        norm("duration_predictor.norm_2", "dp.norm_2")
        conv("duration_predictor.proj", "dp.proj")
    if multi:
        conv("duration_predictor.cond", "dp.cond")
        conv("decoder.cond", "dec.cond")
        put("embed_speaker.weight", params["emb_g.weight"])
    for n in range(sv.FLOW_N_FLOWS):
        mine, hf = f"flow.flows.{2 * n}", f"flow.flows.{n}"
        conv(f"{hf}.conv_pre", f"{mine}.pre")
        conv(f"{hf}.conv_post", f"{mine}.post")
        for i in range(sv.FLOW_WN_LAYERS):
            wn_conv(f"{hf}.wavenet.in_layers.{i}", f"{mine}.enc.in_layers.{i}")
            wn_conv(f"{hf}.wavenet.res_skip_layers.{i}", f"{mine}.enc.res_skip_layers.{i}")
        if multi:
            wn_conv(f"{hf}.wavenet.cond_layer", f"{mine}.enc.cond_layer")
    conv("decoder.conv_pre", "dec.conv_pre")
    conv("decoder.conv_post", "dec.conv_post")
    nk = len(cfg.resblock_kernel_sizes)
    for i in range(len(cfg.upsample_rates)):
        conv(f"decoder.upsampler.{i}", f"dec.ups.{i}")
        for j in range(nk):
            for d in range(len(cfg.resblock_dilation_sizes[j])):
                if rb2:
                    conv(f"decoder.resblocks.{i * nk + j}.convs1.{d}", f"dec.resblocks.{i * nk + j}.convs.{d}")
                    continue
                conv(f"decoder.resblocks.{i * nk + j}.convs1.{d}", f"dec.resblocks.{i * nk + j}.convs1.{d}")
                conv(f"decoder.resblocks.{i * nk + j}.convs2.{d}", f"dec.resblocks.{i * nk + j}.convs2.{d}")
    missing = [k for k in sd if k not in used and not k.startswith(("posterior_encoder.", "duration_predictor.post_"))
               and not k.startswith("duration_predictor.flows.1.")]
    assert not missing, missing[:8]
    model.load_state_dict(sd)
    return model


@pytest.mark.parametrize("use_sdp,n_speakers", [(True, 3), (False, 1), (True, 1)])
def test_oracle_matches_transformers_vits(tmp_path, use_sdp, n_speakers):
    torch = pytest.importorskip("torch")
    pytest.importorskip("transformers")
    from oracle.vits_oracle import VitsOracle

    cfg = sv.tiny_config(n_speakers=n_speakers, resblock="1", use_sdp=use_sdp)
    params = sv.write_voice(tmp_path / "v", cfg, seed=5)
    orc = VitsOracle(str(tmp_path / "v"))
    hf = _hf_model(cfg, params)
    rng = np.random.default_rng(17)
    worst = 0.0
    for T, length_scale, sid in ((23, 1.0, 0), (7, 1.3, n_speakers - 1), (40, 0.8, 0), (1, 1.0, 0)):
        ids = rng.integers(4, cfg.num_symbols, size=T).astype(np.int64)
        audio, inter = orc.infer(ids, (0.0, length_scale, 0.0), sid=sid if n_speakers > 1 else None,
                                 return_intermediates=True)
        hf.noise_scale, hf.noise_scale_duration, hf.speaking_rate = 0.0, 0.0, 1.0 / length_scale
        with torch.no_grad():
            out = hf(input_ids=torch.from_numpy(ids)[None], attention_mask=torch.ones(1, T, dtype=torch.long),
                     speaker_id=sid if n_speakers > 1 else None)
        want = out.waveform[0].numpy()
        assert int(out.sequence_lengths[0]) == audio.shape[0] == want.shape[0], (T, out.sequence_lengths, audio.shape)
        rms = float(np.sqrt(np.mean((audio - want) ** 2)))
        sig = float(np.sqrt(np.mean(want ** 2)))
        worst = max(worst, rms)
        assert rms <= 2e-5 * max(1.0, sig / 0.1), (T, rms, sig)
    print(f"sdp={use_sdp} speakers={n_speakers}: worst RMS oracle vs transformers VITS {worst:.2e}")


def test_oracle_matches_transformers_vits_with_noise(tmp_path):
    """Noise scales > 0 (the voice defaults 0.667 / 0.8): both implementations are fed the SAME normal samples
    (the oracle's Philox streams are injected where transformers calls torch.randn / randn_like), so the
    stochastic duration predictor's spline flows and the prior sampling are compared away from z = 0."""
    torch = pytest.importorskip("torch")
    pytest.importorskip("transformers")
    from unittest import mock
    from oracle import philox
    from oracle.vits_oracle import VitsOracle

    cfg = sv.tiny_config(n_speakers=3, resblock="1", use_sdp=True)
    params = sv.write_voice(tmp_path / "v", cfg, seed=9)
    orc = VitsOracle(str(tmp_path / "v"))
    hf = _hf_model(cfg, params)
    rng = np.random.default_rng(3)
    for T, scales, sid, seed in ((31, (0.667, 1.0, 0.8), 1, 7), (12, (0.3, 1.2, 1.0), 2, 8)):
        ids = rng.integers(4, cfg.num_symbols, size=T).astype(np.int64)
        audio = orc.infer(ids, scales, sid=sid, seed=seed, row=0)

        def fake_randn(*size, **kw):
            size = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
            assert size == (1, 2, T), size
            return torch.from_numpy(philox.normal(seed, 0, 0, np.arange(T)[None, :], np.arange(2)[:, None]))[None].float()

        def fake_randn_like(t, **kw):
            assert t.shape[0] == 1 and t.shape[1] == cfg.inter_channels
            fr = t.shape[2]
            return torch.from_numpy(philox.normal(seed, 1, 0, np.arange(fr)[None, :], np.arange(t.shape[1])[:, None]))[None].float()

        hf.noise_scale, hf.speaking_rate, hf.noise_scale_duration = scales[0], 1.0 / scales[1], scales[2]
        with torch.no_grad(), mock.patch("torch.randn", fake_randn), mock.patch("torch.randn_like", fake_randn_like):
            out = hf(input_ids=torch.from_numpy(ids)[None], attention_mask=torch.ones(1, T, dtype=torch.long), speaker_id=sid)
        want = out.waveform[0].numpy()
        assert want.shape == audio.shape
        rms = float(np.sqrt(np.mean((audio - want) ** 2)))
        print(f"T={T} scales={scales}: RMS oracle vs transformers VITS (shared noise) {rms:.2e}, {audio.shape[0]} samples")
        assert rms <= 5e-5


def test_oracle_matches_transformers_vits_at_low_voice_shapes(tmp_path):
    """Same check at the dimensions of the shipped ``*_low`` voices (hidden 192, 2 heads x 96, 6 layers, FFN 768,
    109 speakers, upsampling 8/8/4) with a ResBlock1 generator of the same kernel sizes / dilations."""
    torch = pytest.importorskip("torch")
    pytest.importorskip("transformers")
    from dataclasses import replace
    from oracle.vits_oracle import VitsOracle

    cfg = replace(sv.low_config(n_speakers=109), resblock="1")
    params = sv.write_voice(tmp_path / "v", cfg, seed=31)
    orc = VitsOracle(str(tmp_path / "v"))
    hf = _hf_model(cfg, params)
    ids = np.random.default_rng(5).integers(4, cfg.num_symbols, size=33).astype(np.int64)
    audio = orc.infer(ids, (0.0, 1.0, 0.0), sid=57)
    hf.noise_scale, hf.noise_scale_duration, hf.speaking_rate = 0.0, 0.0, 1.0
    with torch.no_grad():
        out = hf(input_ids=torch.from_numpy(ids)[None], attention_mask=torch.ones(1, 33, dtype=torch.long), speaker_id=57)
    want = out.waveform[0].numpy()
    assert want.shape == audio.shape and audio.shape[0] % 256 == 0
    rms = float(np.sqrt(np.mean((audio - want) ** 2)))
    print(f"low shapes: {audio.shape[0]} samples, signal RMS {np.sqrt(np.mean(want ** 2)):.3f}, oracle vs transformers RMS {rms:.2e}")
    assert rms <= 2e-5


@pytest.mark.parametrize("n_speakers", [3, 1])
def test_oracle_matches_transformers_vits_resblock2(tmp_path, n_speakers):
    """The generator variant every shipped ``*_low`` voice uses (ResBlock2, config.py:128-129 ``resblock: "2"``),
    anchored on HF's unmodified residual-block forward with the second conv replaced by lrelu^-1 (module docstring)."""
    torch = pytest.importorskip("torch")
    pytest.importorskip("transformers")
    from oracle.vits_oracle import VitsOracle

    cfg = sv.tiny_config(n_speakers=n_speakers, resblock="2", use_sdp=True)
    params = sv.write_voice(tmp_path / "v", cfg, seed=6)
    assert any(".convs.0.weight" in k for k in params) and not any(".convs1." in k for k in params)
    orc = VitsOracle(str(tmp_path / "v"))
    hf = _hf_model(cfg, params)
    rng = np.random.default_rng(19)
    for T, length_scale, sid in ((23, 1.0, 0), (7, 1.3, n_speakers - 1), (40, 0.8, 0), (1, 1.0, 0)):
        ids = rng.integers(4, cfg.num_symbols, size=T).astype(np.int64)
        audio = orc.infer(ids, (0.0, length_scale, 0.0), sid=sid if n_speakers > 1 else None)
        hf.noise_scale, hf.noise_scale_duration, hf.speaking_rate = 0.0, 0.0, 1.0 / length_scale
        with torch.no_grad():
            out = hf(input_ids=torch.from_numpy(ids)[None], attention_mask=torch.ones(1, T, dtype=torch.long),
                     speaker_id=sid if n_speakers > 1 else None)
        want = out.waveform[0].numpy()
        assert int(out.sequence_lengths[0]) == audio.shape[0] == want.shape[0]
        rms = float(np.sqrt(np.mean((audio - want) ** 2)))
        sig = float(np.sqrt(np.mean(want ** 2)))
        assert rms <= 2e-5 * max(1.0, sig / 0.1), (T, rms, sig)


def test_oracle_matches_transformers_vits_resblock2_at_low_voice_shapes(tmp_path):
    """ResBlock2 at the exact shapes of the benchmarked voice (`sv.low_config(n_speakers=109)`, the configuration
    bench.py and the GPU parity tests use): the oracle the CUDA path is compared with is itself compared with HF."""
    torch = pytest.importorskip("torch")
    pytest.importorskip("transformers")
    from oracle.vits_oracle import VitsOracle

    cfg = sv.low_config(n_speakers=109)
    assert str(cfg.resblock) == "2"
    params = sv.write_voice(tmp_path / "v", cfg, seed=22)
    orc = VitsOracle(str(tmp_path / "v"))
    hf = _hf_model(cfg, params)
    ids = np.random.default_rng(5).integers(4, cfg.num_symbols, size=40).astype(np.int64)
    for sid in (0, 108):
        audio = orc.infer(ids, (0.0, 1.0, 0.0), sid=sid)
        hf.noise_scale, hf.noise_scale_duration, hf.speaking_rate = 0.0, 0.0, 1.0
        with torch.no_grad():
            out = hf(input_ids=torch.from_numpy(ids)[None], attention_mask=torch.ones(1, 40, dtype=torch.long), speaker_id=sid)
        want = out.waveform[0].numpy()
        assert want.shape == audio.shape and audio.shape[0] % 256 == 0
        rms = float(np.sqrt(np.mean((audio - want) ** 2)))
        print(f"low shapes, ResBlock2, sid {sid}: {audio.shape[0]} samples, oracle vs transformers RMS {rms:.2e}")
        assert rms <= 2e-5
