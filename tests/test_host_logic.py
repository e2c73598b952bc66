"""Host-side logic of the C-ABI library that needs no device: status strings (cv::Error codes), the panorama
geometry of the hand-written linear blend ([BLEND]:141-176) against the oracle, and NULL-argument behaviour."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    from imagestitch_b200 import build as B, capi
    B.build()
    return capi.load()


def test_version_and_status_strings(lib):
    from imagestitch_b200 import capi
    assert lib.is_version().decode().count(".") >= 1
    seen = set()
    for code in (capi.IS_OK, capi.IS_ERR_NO_MEM, capi.IS_ERR_BAD_ARG, capi.IS_ERR_UNSUPPORTED, capi.IS_ERR_ASSERT, capi.IS_ERR_CUDA,
                 capi.IS_ERR_INTERNAL):
        s = lib.is_status_string(code).decode()
        assert s and s not in seen
        seen.add(s)
    # the codes follow cv::Error so that a maintainer can map them onto cv::Exception one to one
    assert (capi.IS_ERR_NO_MEM, capi.IS_ERR_BAD_ARG, capi.IS_ERR_UNSUPPORTED, capi.IS_ERR_ASSERT) == (-4, -5, -213, -215)
    assert lib.is_status_string(12345).decode()          # unknown codes still give a printable string


@pytest.mark.parametrize("case", [((120, 160), (118, 150), (0, 3), (101, 0)), ((64, 64), (64, 64), (5, 5), (40, 9)),
                                  ((200, 90), (180, 120), (-7, 2), (50, -4)), ((33, 47), (35, 41), (0, 0), (46, 1))])
def test_linear_blend_geometry_matches_oracle(lib, case):
    import oracle as O
    from imagestitch_b200 import capi
    O.build()
    (h1, w1), (h2, w2), tl1, tl2 = case
    he, br = C.c_int(0), C.c_int(0)
    O.lib().orc_lin_geometry(C.c_int(h1), C.c_int(w1), C.c_int(h2), C.c_int(w2), C.c_int(tl1[0]), C.c_int(tl1[1]), C.c_int(tl2[0]),
                             C.c_int(tl2[1]), C.byref(he), C.byref(br))
    sz = capi.Size(0, 0)
    assert lib.is_linear_blend_size(capi.Size(w1, h1), capi.Size(w2, h2), capi.Point(*tl1), capi.Point(*tl2), C.byref(sz)) == capi.IS_OK
    assert (sz.height, sz.width) == (he.value, br.value)


def test_null_arguments_are_rejected_without_a_device(lib):
    from imagestitch_b200 import capi
    assert lib.is_linear_blend_size(capi.Size(4, 4), capi.Size(4, 4), capi.Point(0, 0), capi.Point(2, 0), None) == capi.IS_ERR_BAD_ARG
    assert lib.is_ctx_create(0, None) == capi.IS_ERR_BAD_ARG
    m = capi.Mat()
    assert lib.is_warp_roi(None, 0, capi.Size(8, 8), None, None, C.c_float(1.0), None, None) == capi.IS_ERR_BAD_ARG
    assert lib.is_seam_dp_find(None, 0, None, None, None, 0) == capi.IS_ERR_BAD_ARG
    assert lib.is_blender_create(None, 5, capi.IS_32F, None) == capi.IS_ERR_BAD_ARG
    assert lib.is_gain_apply(None, C.byref(m), C.c_double(1.0)) == capi.IS_ERR_BAD_ARG
    assert lib.is_pipeline_run(None, 0, None, None, None, None, None, None, None) == capi.IS_ERR_BAD_ARG
