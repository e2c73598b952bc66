"""Identities the tiled level-0 blend kernel (csrc/blend.cu, k_blend_l0_tiled) relies on, checked exhaustively in float32."""
import numpy as np


def test_mask255_weight_is_exactly_one():
    inv = np.float32(1.0 / 255.0)
    assert np.float32(255.0) * inv == np.float32(1.0)
    # every other mask value gives a weight strictly inside (0, 1)
    m = np.arange(1, 255, dtype=np.float32)
    w = m * inv
    assert np.all(w > 0) and np.all(w < 1)


def test_normalise_with_unit_weight_is_d_minus_sign():
    # MultiBandBlender::blend: dst = short(dst / (w + 1e-5f)) with w == 1.0f, for every int16 dst
    d = np.arange(-32768, 32768, dtype=np.int32)
    den = np.float32(1.0) + np.float32(1e-5)
    q = (d.astype(np.float32) / den).astype(np.float32)
    got = np.trunc(q).astype(np.int32)
    want = d - np.sign(d)
    assert np.array_equal(got, want)


def test_pyrup_shift_folding():
    # (4 a + 32) >> 6 == (a + 8) >> 4 and (16 a + 32) >> 6 == (a + 2) >> 2 for the ranges pyrUp produces (floor shifts)
    a = np.arange(-8 * 32768 * 8, 8 * 32768 * 8, 37, dtype=np.int64)
    assert np.array_equal((4 * a + 32) >> 6, (a + 8) >> 4)
    assert np.array_equal((16 * a + 32) >> 6, (a + 2) >> 2)


def test_pyrup_of_int16_stays_in_int16():
    # pyrUp outputs are rounded weighted means with weights summing to 64 (k_blend_level_quad drops the saturation there)
    rng = np.random.default_rng(0)
    v = rng.integers(-32768, 32768, (200000, 3, 3)).astype(np.int64)
    v[:4] = np.asarray([-32768, 32767, -32768, 32767]).reshape(4, 1, 1)
    e = v[:, :, 0] + 6 * v[:, :, 1] + v[:, :, 2]
    o = 4 * (v[:, :, 1] + v[:, :, 2])
    outs = [(e[:, 0] + 6 * e[:, 1] + e[:, 2] + 32) >> 6, (o[:, 0] + 6 * o[:, 1] + o[:, 2] + 32) >> 6, (4 * (e[:, 1] + e[:, 2]) + 32) >> 6, (4 * (o[:, 1] + o[:, 2]) + 32) >> 6]
    for x in outs:
        assert x.min() >= -32768 and x.max() <= 32767
