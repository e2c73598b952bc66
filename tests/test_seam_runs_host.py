"""Host logic of the batched seam path (csrc/seam_runs.inl: PairRuns) without a GPU: the run-domain structure of an image
pair -- components and their numbering ([SEAM]:196-308), contour records, conflict-loop plan and seam tips ([SEAM]:395-546,
607-706) -- through the diagnostic entry is_debug_seam_pair_plan, against (a) a brute-force numpy restatement of the
definitions on the pixel grid and (b) the seams the oracle traces for the same pair (== the reference's own find())."""
import ctypes as C

import numpy as np
import pytest
from scipy import ndimage

from helpers import blob_masks, seam_edge_cases, warped_set
from imagestitch_b200 import capi

ST_FIRST, ST_SECOND, ST_INTERS = 1, 2, 4


def plan(m1, tl1, m2, tl2):
    lib = capi.load()
    m1 = np.ascontiguousarray(m1, np.uint8)
    m2 = np.ascontiguousarray(m2, np.uint8)
    n = C.c_size_t(0)
    args = (m1.ctypes.data, m1.shape[0], m1.shape[1], m1.strides[0], tl1[0], tl1[1], m2.ctypes.data, m2.shape[0], m2.shape[1], m2.strides[0], tl2[0], tl2[1])
    assert lib.is_debug_seam_pair_plan(*args, None, 0, C.byref(n)) == 0
    out = np.zeros(n.value, np.int32)
    assert lib.is_debug_seam_pair_plan(*args, out.ctypes.data_as(C.POINTER(C.c_int32)), out.size, C.byref(n)) == 0
    too_many, unsupported, ncomps, nops, nrec, ux, uy = (int(v) for v in out[:7])
    k = 7
    states = out[k:k + ncomps].copy(); k += ncomps
    ops = out[k:k + 11 * nops].reshape(nops, 11).copy(); k += 11 * nops
    recs = out[k:k + 7 * nrec].reshape(nrec, 7).copy()
    return dict(too_many=too_many, unsupported=unsupported, ncomps=ncomps, states=states, ops=ops, recs=recs, union_tl=(ux, uy))


def brute(m1, tl1, m2, tl2):
    """labels / states / contour records of the INTERS components from the pixel grid"""
    ux, uy = min(tl1[0], tl2[0]), min(tl1[1], tl2[1])
    bx = max(tl1[0] + m1.shape[1], tl2[0] + m2.shape[1]); by = max(tl1[1] + m1.shape[0], tl2[1] + m2.shape[0])
    uw, uh = bx - ux, by - uy
    cls = np.zeros((uh, uw), np.int32)
    cls[tl1[1] - uy:tl1[1] - uy + m1.shape[0], tl1[0] - ux:tl1[0] - ux + m1.shape[1]] |= (m1 != 0).astype(np.int32)
    cls[tl2[1] - uy:tl2[1] - uy + m2.shape[0], tl2[0] - ux:tl2[0] - ux + m2.shape[1]] |= (m2 != 0).astype(np.int32) * 2
    firsts = []
    lab_c = {}
    for c in (1, 2, 3):
        lab, k = ndimage.label(cls == c)             # 4-connectivity
        lab_c[c] = lab
        if k:
            idx = ndimage.minimum(np.arange(uh * uw).reshape(uh, uw), lab, np.arange(1, k + 1))
            firsts += [(int(i), c, j + 1) for j, i in enumerate(np.atleast_1d(idx))]
    firsts.sort()
    labels = np.zeros((uh, uw), np.int32)
    states = []
    for new, (_, c, old) in enumerate(firsts):
        labels[lab_c[c] == old] = new + 1
        states.append({1: ST_FIRST, 2: ST_SECOND, 3: ST_INTERS}[c])
    pad = np.full((uh + 2, uw + 2), -1, np.int32)
    pad[1:-1, 1:-1] = labels
    recs = []
    for l, st in enumerate(states, 1):
        if st != ST_INTERS:
            continue
        ys, xs = np.nonzero(labels == l)
        for y, x in zip(ys, xs):
            nl = [pad[y + 1, x], pad[y, x + 1], pad[y + 1, x + 2], pad[y + 2, x + 1]]      # left, up, right, down
            if any(v != l for v in nl):
                recs.append([x, y, l] + [int(v) for v in nl])
    return dict(ncomps=len(states), states=np.asarray(states, np.int32), recs=np.asarray(recs, np.int32).reshape(-1, 7), union_tl=(ux, uy))


def _pairs():
    import oracle as O
    O.build()
    cases = []
    for (n, w, h, ov, rows) in ((2, 260, 200, 0.25, 1), (3, 200, 150, 0.6, 1), (4, 160, 120, 0.3, 2), (6, 120, 100, 0.3, 2)):
        corners, wi, wm = warped_set(O, n, w, h, overlap=ov, grid_rows=rows)
        for i in range(n):
            for j in range(i + 1, n):
                cases.append((f"warped n={n} rows={rows} ({i},{j})", wi[i], wi[j], corners[i], corners[j], wm[i], wm[j]))
    corners, wi, wm = warped_set(O, 3, 180, 130, overlap=0.4)
    holes = blob_masks(np.random.default_rng(8), [m.shape for m in wm], holes=4)
    wm = [np.where(hm > 0, m, 0).astype(np.uint8) for m, hm in zip(wm, holes)]
    for (i, j) in ((0, 1), (1, 2), (0, 2)):
        cases.append((f"holes ({i},{j})", wi[i], wi[j], corners[i], corners[j], wm[i], wm[j]))
    for name, imgs, cs, ms, cost in seam_edge_cases():
        if cost == 0:
            cases.append((name + " (0,1)", imgs[0].astype(np.uint8), imgs[1].astype(np.uint8), cs[0], cs[1], ms[0], ms[1]))
    # random rectangles with a notch or a hole each, random placement (horizontal and vertical seams, components that split)
    for seed in range(40):
        r = np.random.default_rng(1000 + seed)
        h1, w1, h2, w2 = (int(v) for v in r.integers(40, 110, 4))
        a = r.integers(0, 256, (h1, w1, 3)).astype(np.uint8)
        b = r.integers(0, 256, (h2, w2, 3)).astype(np.uint8)
        m1, m2 = blob_masks(r, [(h1, w1), (h2, w2)], holes=int(r.integers(0, 3)))
        c2 = (int(r.integers(-w2 + 8, w1 - 8)), int(r.integers(-h2 + 8, h1 - 8)))
        cases.append((f"random {seed}", a, b, (0, 0), c2, m1, m2))
    return cases


_CASES = None


def _cases():
    global _CASES
    if _CASES is None:
        _CASES = _pairs()
    return _CASES


def test_pair_structure_and_plan_against_grid_and_oracle(oracle):
    O = oracle
    checked_plans = 0
    for name, a, b, c1, c2, m1, m2 in _cases():
        c1 = (int(c1[0]), int(c1[1])); c2 = (int(c2[0]), int(c2[1]))
        overlap = max(c1[0], c2[0]) < min(c1[0] + m1.shape[1], c2[0] + m2.shape[1]) and max(c1[1], c2[1]) < min(c1[1] + m1.shape[0], c2[1] + m2.shape[0])
        if not overlap:
            continue
        got = plan(m1, c1, m2, c2)
        if got["too_many"]:
            continue                                   # the device path hands these to the general path
        want = brute(m1, c1, m2, c2)
        assert got["ncomps"] == want["ncomps"], name
        assert np.array_equal(got["states"], want["states"]), name
        assert got["recs"].shape == want["recs"].shape and np.array_equal(got["recs"], want["recs"]), f"{name}: contour records"
        if got["unsupported"]:                         # 1: not covered; 2: staged plan (only its first round is listed here)
            continue
        # the seams the reference estimates for this pair: component and end points = the plan's seam operations
        _, trace = O.dp_seam_find([a, b], [c1, c2], [m1, m2], want_trace=True)
        dp_ops = [op for op in got["ops"] if op[0] == 1]
        ux, uy = got["union_tl"]
        # estimateSeam can fail (destination unreachable): such seams are planned but leave no trace; match in order
        k = 0
        for (_, _, comp, horiz, pts) in trace:
            while k < len(dp_ops) and not (dp_ops[k][1] == comp and tuple(pts[0]) == (dp_ops[k][3] + ux, dp_ops[k][4] + uy)
                                           and tuple(pts[-1]) == (dp_ops[k][5] + ux, dp_ops[k][6] + uy)):
                k += 1
            assert k < len(dp_ops), f"{name}: the oracle's seam of component {comp} {tuple(pts[0])}->{tuple(pts[-1])} is not in the plan {dp_ops}"
            k += 1
            checked_plans += 1
    assert checked_plans >= 20


def test_pair_finish_on_host_equals_oracle(oracle):
    """The whole pair on the CPU through the batched path's host code -- plan, run-domain updateLabelsUsingSeam, clear intervals --
    fed with the seams the oracle traces: the resulting masks equal the oracle's (== the reference's own find())."""
    O = oracle
    lib = capi.load()
    done = staged = staged_done = 0
    for name, a, b, c1, c2, m1, m2 in _cases():
        c1 = (int(c1[0]), int(c1[1])); c2 = (int(c2[0]), int(c2[1]))
        got = plan(m1, c1, m2, c2)
        if got["too_many"] or got["unsupported"] == 1:
            continue
        want, trace = O.dp_seam_find([a, b], [c1, c2], [m1, m2], want_trace=True)
        ux, uy = got["union_tl"]
        dp_ops = [op for op in got["ops"] if op[0] == 1]
        seams, k = [], 0
        for op in dp_ops:                              # the oracle's seams in the plan's order; a seam that failed leaves no trace
            p1 = (op[3] + ux, op[4] + uy); p2 = (op[5] + ux, op[6] + uy)
            if k < len(trace) and trace[k][2] == op[1] and tuple(trace[k][4][0]) == p1 and tuple(trace[k][4][-1]) == p2:
                pts = trace[k][4]
                seams += [len(pts)] + [int(v) for v in pts.reshape(-1)]
                k += 1
            else:
                seams.append(0)
        if got["unsupported"] == 2:
            # staged plan: only the first round is listed; the later rounds take the oracle's remaining seams in order (an operation
            # whose estimateSeam failed cannot be told apart here: such cases are skipped)
            staged += 1
            for t in trace[k:]:
                seams += [len(t[4])] + [int(v) for v in t[4].reshape(-1)]
            k = len(trace)
        assert k == len(trace), name
        o1 = np.ascontiguousarray(m1).copy(); o2 = np.ascontiguousarray(m2).copy()
        sa = np.asarray(seams if seams else [0], np.int32)
        rc = lib.is_debug_seam_pair_finish(o1.ctypes.data, o1.shape[0], o1.shape[1], o1.strides[0], c1[0], c1[1], o2.ctypes.data, o2.shape[0], o2.shape[1],
                                           o2.strides[0], c2[0], c2[1], sa.ctypes.data_as(C.POINTER(C.c_int32)), len(seams))
        if got["unsupported"] == 2 and rc != 0:
            continue                                   # a later round met what the plan does not cover (or a failed estimateSeam shifted the seam list)
        assert rc == 0, (name, rc)
        assert np.array_equal(o1, want[0]) and np.array_equal(o2, want[1]), f"{name}: masks differ in {int((o1 != want[0]).sum())} + {int((o2 != want[1]).sum())} px"
        done += 1
        staged_done += 1 if got["unsupported"] == 2 else 0
    assert done >= 25
    print(f"staged pairs: {staged_done} of {staged} finished on the host")


def _wave_schedule(corners, sizes):
    lib = capi.load()
    n = len(corners)
    cs = (capi.Point * n)(*[capi.Point(int(c[0]), int(c[1])) for c in corners])
    ss = (capi.Size * n)(*[capi.Size(int(s[0]), int(s[1])) for s in sizes])
    need = C.c_size_t(0)
    assert lib.is_debug_seam_wave_schedule(n, cs, ss, None, 0, C.byref(need)) == 0
    out = np.zeros(need.value, np.int32)
    assert lib.is_debug_seam_wave_schedule(n, cs, ss, out.ctypes.data_as(C.POINTER(C.c_int32)), out.size, C.byref(need)) == 0
    waves, k = [], 1
    for _ in range(int(out[0])):
        cnt = int(out[k]); k += 1
        waves.append([(int(out[k + 2 * q]), int(out[k + 2 * q + 1])) for q in range(cnt)])
        k += 2 * cnt
    return waves


@pytest.mark.parametrize("case", [(6, 200, 150, 1, 1.2, 0.25), (12, 200, 140, 4, 1.5, 0.25), (12, 180, 140, 3, 1.3, 0.3), (16, 150, 110, 4, 1.4, 0.3), (10, 220, 160, 2, 1.2, 0.35)])
def test_wave_order_commutes_with_the_reference_order(oracle, case):
    """The batched path lets pairs overtake pairs they share no image with.  Run through the oracle one pair at a time, the wave
    order must leave exactly the masks of the reference's order ([SEAM]:100-121); the schedule itself: every overlapping pair once,
    the first remaining pair always in the wave, no wave member shares an image with a pair that was left out in front of it,
    a strip is a single wave."""
    O = oracle
    n, w, h, rows, fw, ov = case
    corners, wi, wm = warped_set(O, n, w, h, f_over_w=fw, overlap=ov, grid_rows=rows)
    corners = [(int(c[0]), int(c[1])) for c in corners]
    sizes = [(m.shape[1], m.shape[0]) for m in wm]
    waves = _wave_schedule(corners, sizes)
    ref_order = [(i, j) for i in range(n - 1) for j in range(i + 1, n)][::-1]
    overlapping = [(i, j) for (i, j) in ref_order
                   if max(corners[i][0], corners[j][0]) < min(corners[i][0] + sizes[i][0], corners[j][0] + sizes[j][0])
                   and max(corners[i][1], corners[j][1]) < min(corners[i][1] + sizes[i][1], corners[j][1] + sizes[j][1])]
    flat = [p for wv in waves for p in wv]
    assert sorted(flat) == sorted(overlapping) and len(set(flat)) == len(flat)
    if rows == 1:
        assert len(waves) == 1 and waves[0] == overlapping
    done = set()
    for wv in waves:
        remaining = [p for p in overlapping if p not in done]
        assert wv[0] == remaining[0]
        pos = {p: k for k, p in enumerate(remaining)}
        assert [pos[p] for p in wv] == sorted(pos[p] for p in wv), "wave members keep the reference's order among themselves"
        for p in wv:
            skipped_before = [q for q in remaining[:pos[p]] if q not in wv]
            assert not any(set(p) & set(q) for q in skipped_before), f"{p} overtakes a pair it shares an image with"
        done.update(wv)
    # the reference's loop vs the wave order, one pair at a time through the oracle
    def run(order):
        masks = [m.copy() for m in wm]
        for (i, j) in order:
            out = O.dp_seam_find([wi[i], wi[j]], [corners[i], corners[j]], [masks[i], masks[j]])
            masks[i], masks[j] = out[0], out[1]
        return masks
    a, b = run(overlapping), run(flat)
    for k in range(n):
        assert np.array_equal(a[k], b[k]), f"mask {k} differs between the reference's order and the wave order"
    if rows > 1:
        assert len(waves) < len(overlapping)
