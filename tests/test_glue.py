"""Oracle tier A: the reference's in-tree arithmetic (av1encoder.rs:485-605, dirtyalpha.rs) pinned by the
facts of SURVEY.md App. A and the reference's own known-answer test (dirtyalpha.rs:126-135); and the
product's host-side restatement (ravif_settings.cpp) checked bit-exactly against the oracle."""
import ctypes
import numpy as np
import pytest
import harness as H


def _orc():
    L = H.oracle()
    L.orc_quality_to_quantizer.restype = ctypes.c_uint8
    L.orc_quality_to_quantizer.argtypes = [ctypes.c_float]
    L.orc_tile_target.restype = ctypes.c_size_t
    L.orc_tile_target.argtypes = [ctypes.c_size_t] * 4
    return L


def test_quality_to_quantizer_table():
    L = _orc()
    exp = {1: 252, 22: 199, 25: 191, 33: 181, 55: 153, 60: 147, 66: 139, 77: 125, 80: 121, 82: 119, 88: 80, 90: 66, 95: 33, 100: 0}
    P = H.hostlib()
    P.b200avif_quality_to_quantizer.restype = ctypes.c_uint8
    P.b200avif_quality_to_quantizer.argtypes = [ctypes.c_float]
    for q, v in exp.items():
        assert L.orc_quality_to_quantizer(q) == v
    for q in np.arange(1, 100.01, 0.25):
        assert L.orc_quality_to_quantizer(float(q)) == P.b200avif_quality_to_quantizer(float(q))


def test_preminmax_known_answers():      # dirtyalpha.rs:126-135
    L = _orc()
    o = (ctypes.c_uint8 * 2)()
    for (px, a), exp in {(100, 255): (100, 100), (100, 10): (78, 100), (100, 2): (8, 119), (100, 1): (16, 239), (255, 1): (15, 255)}.items():
        L.orc_premultiplied_minmax(px, a, o)
        assert tuple(o) == exp


def test_colour_facts():
    L = _orc()
    o = (ctypes.c_uint16 * 3)()
    cases = {((255, 255, 255), 8): (255, 128, 128), ((255, 255, 255), 10): (1023, 512, 512), ((0, 0, 0), 8): (0, 128, 128),
             ((0, 0, 0), 10): (0, 512, 512), ((128, 128, 128), 8): (128, 128, 128), ((128, 128, 128), 10): (514, 512, 512)}
    for (rgb, d), exp in cases.items():
        L.orc_pixel_planes(*rgb, 0, d, o)
        assert tuple(o) == exp
    L.orc_pixel_planes(255, 0, 0, 0, 8, o); assert o[2] == 255       # 256 saturated by `as u8`
    L.orc_pixel_planes(0, 0, 255, 0, 8, o); assert o[1] == 255
    L.orc_pixel_planes(255, 0, 0, 0, 10, o); assert o[2] == 1024     # `as u16` does not clamp to 1023
    L.orc_pixel_planes(10, 20, 30, 1, 8, o); assert tuple(o) == (20, 30, 10)   # G,B,R
    L.orc_pixel_planes(255, 0, 128, 1, 10, o); assert tuple(o) == (0, 514, 1023)


def test_speed_tweaks_table():
    class T(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int) for n in ("speed_preset", "fast_deblock", "reduced_tx_set", "tx_domain_distortion", "tx_domain_rate",
                                                "encode_bottomup", "rdo_tx_decision", "cdef", "lrf", "sgr_complexity_full", "use_satd_subpel",
                                                "inter_tx_split", "fine_directional_intra", "complex_prediction_modes", "partition_min",
                                                "partition_max", "min_tile_size")]
    class P(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int) for n in ("speed_preset", "fast_deblock", "reduced_tx_set", "tx_domain_rate", "encode_bottomup",
                                                "rdo_tx_decision", "cdef", "lrf", "sgr_complexity_full", "inter_tx_split",
                                                "fine_directional_intra", "complex_prediction_modes", "partition_min", "partition_max",
                                                "min_tile_size")]
    L = _orc(); Ph = H.hostlib()
    rows = {(4, 121): dict(partition_min=4, partition_max=16, encode_bottomup=0, rdo_tx_decision=1, reduced_tx_set=1, cdef=1, lrf=1, min_tile_size=256),
            (6, 147): dict(partition_min=8, partition_max=16, rdo_tx_decision=0, reduced_tx_set=0, cdef=1, lrf=1, min_tile_size=256),
            (6, 121): dict(partition_min=8, partition_max=16, min_tile_size=128),
            (2, 66): dict(partition_min=4, partition_max=32, encode_bottomup=1, rdo_tx_decision=1, reduced_tx_set=0, sgr_complexity_full=1, min_tile_size=1024)}
    for (s, q), exp in rows.items():
        t = T(); L.orc_speed_tweaks_from_my_preset(s, q, ctypes.byref(t))
        for k, v in exp.items():
            assert getattr(t, k) == v, (s, q, k)
    for s in range(0, 11):
        for q in range(0, 256, 3):
            t = T(); p = P()
            L.orc_speed_tweaks_from_my_preset(s, q, ctypes.byref(t)); Ph.b200_speed_tweaks(s, q, ctypes.byref(p))
            for k, _ in P._fields_:
                assert getattr(t, k) == getattr(p, k), (s, q, k)
    assert L.orc_tile_target(64, 1920, 1080, 256) == 31 and L.orc_tile_target(8, 1920, 1080, 256) == 8
    assert L.orc_tile_target(1000, 3840, 2160, 256) == 126


@pytest.mark.parametrize("seed", [0, 1])
def test_dirty_alpha_oracle_properties(seed):
    """the oracle's dirtyalpha restatement (the device kernels are compared with it in tests/test_gpu_parity.py)"""
    L = _orc()
    img = H.synth_image(97, 61, seed, alpha=True)
    a = np.zeros_like(img)
    assert L.orc_blurred_dirty_alpha(img.ctypes.data, 97, 61, 97, a.ctypes.data) == 1
    assert np.array_equal(a[..., 3], img[..., 3])            # alpha untouched
    opaque = img[..., 3] == 255
    assert np.array_equal(a[opaque], img[opaque])            # opaque pixels untouched
    assert not np.array_equal(a[img[..., 3] == 0], img[img[..., 3] == 0])
    op = img.copy(); op[..., 3] = 255
    assert L.orc_blurred_dirty_alpha(op.ctypes.data, 97, 61, 97, a.ctypes.data) == 0   # None for opaque images
    L.orc_premultiplied_branch(img.ctypes.data, 97, 61, 97, a.ctypes.data)
    semi = (img[..., 3] != 0) & (img[..., 3] != 255)
    assert not a[~semi].any()
