"""is_orb_find (find(image, features) of [FEAT]:948) on the device against the oracle, which tests/test_oracle_orb.py pins to
cv2.ORB: all key-point fields bit for bit, in the reference's order, and every descriptor byte."""
import numpy as np
import pytest

from imagestitch_b200 import stitching as S, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [("synth", (3, 1), 3, (1200, 800)), ("synth", (1, 1), 3, (640, 480)), ("noise", (2, 2), 1, (500, 380)), ("bgra", (3, 1), 4, (900, 600)),
                                  ("noise", (3, 1), 1, (3000, 2000))])
def test_orb_find_matches_oracle(ctx, oracle, case):
    import torch
    O = oracle
    kind, grid, ch, (w, h) = case
    rng = np.random.default_rng(17)
    if kind == "noise":                                    # box-filtered noise: corners at every pyramid level, not only the first
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if w >= 2000:
            b = rng.integers(0, 256, (h // 4 + 1, w // 4 + 1), dtype=np.uint8)
            img = np.ascontiguousarray(np.kron(b, np.ones((4, 4), np.uint8))[:h, :w])
    else:
        img = synth.make_panorama_inputs(2, w, h, 1.2, 0.25)[0][1]
        if kind == "bgra":
            img = np.ascontiguousarray(np.concatenate([img, rng.integers(0, 256, img.shape[:2] + (1,), dtype=np.uint8)], axis=2))
    want_k, want_d = O.orb_find(img, grid)
    launches = ctx.kernel_launches
    got_k, got_d = S.orb_find(ctx, img, grid)
    assert ctx.kernel_launches > launches
    assert len(got_k) == len(want_k) and len(got_k) > 50
    assert np.array_equal(got_k.view(np.uint32), want_k.view(np.uint32)), "key points"
    assert np.array_equal(got_d, want_d), "descriptors"
    dev_k, dev_d = S.orb_find(ctx, torch.from_numpy(img).cuda(), grid)
    assert np.array_equal(dev_k.view(np.uint32), want_k.view(np.uint32)) and np.array_equal(dev_d, want_d), "device-resident image"


def test_orb_rejects_bad_input(ctx):
    with pytest.raises(Exception):
        S.orb_find(ctx, np.zeros((100, 100, 2), np.uint8))
    with pytest.raises(Exception):
        S.orb_find(ctx, np.zeros((100, 100), np.uint8), grid_wh=(8, 8), nlevels=5)
    k, d = S.orb_find(ctx, np.zeros((200, 300), np.uint8))          # nothing to find: zero key points, no error
    assert len(k) == 0 and d.shape == (0, 32)
