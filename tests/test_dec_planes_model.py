"""CPU model of the index arithmetic behind dec_planes_kernel (mimic3_b200/csrc/kernels_tc_dec3.cu): the phase-major plane
layout, the (plane, row shift) of a conv tap, the regrouped conv_post weights of kernels.h and the window origins of its
launcher -- restated in numpy from the comments there and checked against direct convolutions.  (The kernel itself is
checked against the oracle in the -m gpu tests; this pins the mapping it is built on.)"""
import numpy as np

U = 4          # upsampling factor of the last generator stage = number of planes
ROWS = 128     # rows per plane


def to_planes(x, halo):
    """x [512, C] in sample order -> [4][ROWS + 2 halo][C], sample 4t+ph = row halo+t of plane ph, halos zero."""
    p = np.zeros((U, ROWS + 2 * halo, x.shape[1]), x.dtype)
    for ph in range(U):
        p[ph, halo:halo + ROWS] = x[ph::U]
    return p


def conv_on_planes(planes, halo, w, dil):
    """Conv1d (taps x Cin x Cout, 'same' padding, dilation dil) evaluated the way the issuer warps schedule it: for output
    plane ph and tap offset o the operand is plane (ph+o) mod 4 shifted by floor((ph+o)/4) rows."""
    taps = w.shape[0]
    out = np.zeros((U, ROWS, w.shape[2]))
    for ph in range(U):
        for t in range(taps):
            o = (t - (taps - 1) // 2) * dil
            pi, sh = (ph + o) % U, (ph + o) // U          # python floor division == the kernel's (q >> 2) - 64 trick
            assert abs(sh) <= (abs(o) + 3) // 4 <= halo    # plane_halo() of the kernel covers every shift
            out[ph] += planes[pi, halo + sh: halo + sh + ROWS] @ w[t]
    return out


def direct_conv(x, w, dil):
    taps, pad = w.shape[0], (w.shape[0] - 1) // 2 * dil
    xp = np.pad(x, ((pad, pad), (0, 0)))
    return sum(xp[t * dil: t * dil + x.shape[0]] @ w[t] for t in range(taps))


def test_tap_to_plane_and_shift_equals_direct_convolution():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((U * ROWS, 8))
    for taps, dil in ((3, 1), (3, 2), (5, 2), (5, 6), (7, 3), (7, 12), (7, 1)):   # the `_low` generator's convs + conv_post
        w = rng.standard_normal((taps, 8, 5))
        halo = ((taps - 1) // 2 * dil + 3) // 4
        got = conv_on_planes(to_planes(x, halo), halo, w, dil)
        want = direct_conv(x, w, dil)
        for ph in range(U):
            np.testing.assert_allclose(got[ph], want[ph::U], rtol=1e-12, atol=1e-12)


def test_regrouped_conv_post_weights():
    """kernels.h: one [C][16] block per (row shift sh, input plane pi) in the order sh=-1: pi 1..3 | sh=0: pi 0..3 | sh=+1:
    pi 0..2; block column ph' = w[tap 4 sh + pi - ph' + 3].  The four samples 4t..4t+3 of a row are then the four columns of
    sum over blocks of plane_pi[t + sh] @ block."""
    rng = np.random.default_rng(1)
    C = 32
    w = rng.standard_normal((7, C))                 # conv_post: 7 taps, one output channel
    x = rng.standard_normal((U * ROWS, C))
    pairs = [(sh, pi) for sh in (-1, 0, 1) for pi in range(U) if not ((sh < 0 and pi == 0) or (sh > 0 and pi == 3))]
    assert len(pairs) == 10
    blocks = np.zeros((10, C, 16))
    for k, (sh, pi) in enumerate(pairs):
        for php in range(U):
            tap = 4 * sh + pi - php + 3
            if 0 <= tap <= 6:
                blocks[k, :, php] = w[tap]
    used = {4 * sh + pi - php + 3 for sh, pi in pairs for php in range(U)} & set(range(7))
    assert used == set(range(7))
    planes = to_planes(x, 1)
    acc = np.zeros((ROWS, 16))
    for k, (sh, pi) in enumerate(pairs):            # issuer A: blocks 0-4, issuer B: blocks 5-9, summed by the epilogue
        acc += planes[pi, 1 + sh: 1 + sh + ROWS] @ blocks[k]
    want = direct_conv(x, w[:, :, None], 1)[:, 0]
    for php in range(U):
        np.testing.assert_allclose(acc[:, php], want[php::U], rtol=1e-12, atol=1e-12)
    assert not acc[:, U:].any()


def test_polyphase_transposed_conv_lands_in_plane_order():
    """ConvTranspose1d(k = 2u, stride u, padding (k-u)/2): with the window origin chosen so that (w0 + pad) % u == 0, the
    polyphase result D[t][ph] = sum_d y[tq0 + t - d] . W[u d + ph] IS sample w0 + 4t + ph -- plane ph, row t, no lane shift."""
    rng = np.random.default_rng(2)
    u, k = U, 2 * U
    pad = (k - u) // 2
    y = rng.standard_normal(400)
    W = rng.standard_normal(k)
    full = np.zeros(u * len(y) + k)
    for t in range(len(y)):
        full[u * t: u * t + k] += y[t] * W          # out[u t - pad + j] += y[t] W[j]
    x = full[pad: pad + u * len(y)]
    H = 48
    HL = H + ((pad - H) % u + u) % u                # the launcher's left halo
    hr = H
    while (U * ROWS - HL - hr) % u:
        hr += 1
    stride = U * ROWS - HL - hr
    assert (HL, stride) == (50, 412) and stride % u == 0
    for win in (1, 2):
        w0 = win * stride - HL
        assert (w0 + pad) % u == 0
        tq0 = (w0 + pad) // u
        for t in (0, 17, ROWS - 1):
            for ph in range(u):
                d_val = sum(y[tq0 + t - d] * W[u * d + ph] for d in (0, 1))
                np.testing.assert_allclose(d_val, x[w0 + u * t + ph], rtol=1e-12, atol=1e-12)
