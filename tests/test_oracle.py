"""CPU tests of the oracle (test infrastructure) against everything that can pin it here."""
import json
import math
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import onnx_min, philox
from oracle.vits_oracle import VitsOracle, audio_float_to_int16

GOLD = Path(__file__).resolve().parent / "golden"


def test_int16_matches_reference_function_golden():
    """oracle.audio_float_to_int16 == reference mimic3_tts/utils.py:237-244 (fixture made by
    importing the real function, tests/golden/make_golden.py)."""
    g = np.load(GOLD / "int16_reference.npz")
    n = len([k for k in g.files if k.startswith("in")])
    assert n >= 7
    for i in range(n):
        got = audio_float_to_int16(g[f"in{i}"])
        assert got.dtype == np.int16
        np.testing.assert_array_equal(got, g[f"out{i}"])


def test_reference_golden_wav_properties():
    """What the reference's golden WAVs (tests/apope_sample_*.wav) pin at this boundary:
    sample count is a multiple of hop 256 and the peak is exactly 32767."""
    stats = json.loads((GOLD / "golden_wav_stats.json").read_text())
    for arch, s in stats.items():
        assert s["mod_hop256"] == 0 and s["peak"] == 32767 and s["rate"] == 22050
    assert len({s["frames"] for s in stats.values()}) == 1  # identical durations across ISAs


def test_int16_truncates_toward_zero_and_floor():
    x = np.array([0.5, -0.5, 0.505050, -0.2525], dtype=np.float32)
    y = audio_float_to_int16(x)
    assert y[2] == 32767
    s = np.float32(32767.0) / np.float32(0.505050)
    assert y[0] == int(np.float32(0.5) * s) and y[1] == -int(np.float32(0.5) * s)
    quiet = np.full(8, 0.001, dtype=np.float32)  # below the 0.01 floor
    assert audio_float_to_int16(quiet)[0] == int(np.float32(0.001) * (np.float32(32767.0) / np.float32(0.01)))


def test_oracle_regression_fixture(voices):
    g = np.load(GOLD / "oracle_tiny.npz")
    for name in ("tiny", "tiny_ms", "tiny_rb1_dp"):
        o = VitsOracle(str(voices(name)))
        audio, inter = o.infer(g[f"{name}_ids"], (0.0, 1.0, 0.0), sid=1, return_intermediates=True)
        np.testing.assert_array_equal(inter["durations"], g[f"{name}_durations"])
        assert audio.shape == g[f"{name}_audio"].shape
        assert np.sqrt(np.mean((audio - g[f"{name}_audio"]) ** 2)) < 1e-5
        assert np.abs(audio_float_to_int16(audio).astype(int) - g[f"{name}_pcm"].astype(int)).max() <= 1
        noisy = o.infer(g[f"{name}_ids"], (0.667, 1.1, 0.8), sid=1, seed=77, row=0)
        assert noisy.shape == g[f"{name}_noisy_audio"].shape


def test_export_styles_resolve_to_same_parameters(voices):
    a = onnx_min.named_parameters(str(voices("tiny_ms") / "generator.onnx"))
    b = onnx_min.named_parameters(str(voices("tiny_ms_folded") / "generator.onnx"))
    c = onnx_min.named_parameters(str(voices("tiny_ms_wn") / "generator.onnx"))
    keys = [k for k in a if k.endswith(".weight") and ".enc." in k]
    assert keys
    for k in keys:
        np.testing.assert_array_equal(a[k], b[k])
        np.testing.assert_allclose(a[k], c[k], rtol=2e-6, atol=1e-7)


def test_relative_attention_matches_literal_vits_skew(voices):
    """The oracle's gather formulation == VITS' pad/reshape "skewing" formulation."""
    o = VitsOracle(str(voices("tiny")))
    T, W = 13, 4
    nh, dk = o.n_heads, o.H // o.n_heads
    torch.manual_seed(0)
    x = torch.randn(1, o.H, T)
    got = o.attention(x, 0)

    a = "enc_p.encoder.attn_layers.0"
    q = o.conv(x, a + ".conv_q").view(1, nh, dk, T).transpose(2, 3)
    k = o.conv(x, a + ".conv_k").view(1, nh, dk, T).transpose(2, 3)
    v = o.conv(x, a + ".conv_v").view(1, nh, dk, T).transpose(2, 3)

    def rel_emb(e, length):  # attentions._get_relative_embeddings
        pad = max(length - (W + 1), 0)
        s = max((W + 1) - length, 0)
        e = F.pad(e, (0, 0, pad, pad))
        return e[:, s:s + 2 * length - 1]

    def rel_to_abs(x):  # [b,h,l,2l-1] -> [b,h,l,l]
        b, h, l, _ = x.size()
        x = F.pad(x, (0, 1))
        x = x.view(b, h, l * 2 * l)
        x = F.pad(x, (0, l - 1))
        return x.view(b, h, l + 1, 2 * l - 1)[:, :, :l, l - 1:]

    def abs_to_rel(x):  # [b,h,l,l] -> [b,h,l,2l-1]
        b, h, l, _ = x.size()
        x = F.pad(x, (0, l - 1))
        x = x.view(b, h, l ** 2 + l * (l - 1))
        x = F.pad(x, (l, 0))
        return x.view(b, h, l, 2 * l)[:, :, :, 1:]

    qs = q / math.sqrt(dk)
    scores = qs @ k.transpose(-2, -1)
    ek = rel_emb(o.P[a + ".emb_rel_k"], T)
    scores = scores + rel_to_abs(qs @ ek.unsqueeze(0).transpose(-2, -1))
    p = F.softmax(scores, dim=-1)
    out = p @ v
    ev = rel_emb(o.P[a + ".emb_rel_v"], T)
    out = out + abs_to_rel(p) @ ev.unsqueeze(0)
    out = out.transpose(2, 3).contiguous().view(1, o.H, T)
    want = o.conv(out, a + ".conv_o")
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)


def _rqs_forward(x, uw, uh, ud):
    """Forward rational-quadratic spline (Durkan et al.), written independently of the oracle."""
    const = math.log(math.exp(1 - 1e-3) - 1)
    ud = F.pad(ud, (1, 1), value=const)

    def knots(u):
        w = 1e-3 + (1 - 1e-3 * 10) * F.softmax(u, -1)
        c = F.pad(torch.cumsum(w, -1), (1, 0)) * 10 - 5
        c[..., 0], c[..., -1] = -5.0, 5.0
        return c, c[..., 1:] - c[..., :-1]

    cw, w = knots(uw)
    ch, h = knots(uh)
    d = 1e-3 + F.softplus(ud)
    b = (torch.sum(x[..., None] >= cw, -1) - 1).clamp(0, 9)[..., None]
    g = lambda t: t.gather(-1, b)[..., 0]
    theta = (x - g(cw)) / g(w)
    delta = g(h / w)
    d0, d1 = g(d), g(d[..., 1:])
    num = g(h) * (delta * theta ** 2 + d0 * theta * (1 - theta))
    den = delta + (d0 + d1 - 2 * delta) * theta * (1 - theta)
    return g(ch) + num / den


def test_rqs_inverse_inverts_forward():
    torch.manual_seed(1)
    n = 4000
    uw, uh, ud = torch.randn(n, 10) * 1.5, torch.randn(n, 10) * 1.5, torch.randn(n, 9)
    x = (torch.rand(n) * 2 - 1) * 4.99
    y = _rqs_forward(x.double(), uw.double(), uh.double(), ud.double()).float()
    back = VitsOracle.rqs_inverse(y, uw, uh, ud)
    err = torch.abs(back - x)
    assert torch.quantile(err, 0.98) < 1e-3  # fp32 inverse is ill-conditioned only where the slope ~ 0
    again = _rqs_forward(back.double(), uw.double(), uh.double(), ud.double()).float()
    assert torch.max(torch.abs(again - y)) < 2e-4
    far = torch.tensor([-7.0, 6.5])
    assert torch.equal(VitsOracle.rqs_inverse(far, uw[:2], uh[:2], ud[:2]), far)  # linear tails


def test_philox_known_answers_and_moments():
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors): zero and all-ones inputs
    z = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in z] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    o = philox.philox4x32_10(f, f, f, f, f, f)
    assert [int(v) for v in o] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    n = philox.normal(123, 1, 0, np.arange(20000)[None, :], np.arange(4)[:, None])
    assert n.dtype == np.float32 and abs(n.mean()) < 0.02 and abs(n.std() - 1) < 0.02


def test_oracle_edge_cases(voices):
    o = VitsOracle(str(voices("tiny")))
    a1 = o.infer(np.array([5]), (0.0, 1.0, 0.0))  # single phoneme
    hop = int(np.prod(o.m["upsample_rates"]))
    assert a1.size % hop == 0 and a1.size >= hop
    short = o.infer(np.array([5, 6, 7]), (0.0, 1e-6, 0.0))  # every duration ceil()s to 1
    assert short.size == 3 * hop
    a = o.infer(np.arange(4, 16), (0.0, 1.0, 0.0))
    b = o.infer(np.arange(4, 16), (0.0, 2.0, 0.0))
    assert b.size > a.size  # length_scale stretches
