"""Hand-restated normative tables (av1_tables_extra.h) cross-checked by byte-pattern search in the
.rodata of the two independent AV1 implementations on this box (libaom 3.13.1, dav1d 1.5.3)."""
import re, struct, os
import pytest
import aomlib, dav1d
import harness as H


def _rows(name):
    txt = open(os.path.join(H.ROOT, "cavif_rs_b200/csrc/av1_tables_extra.h")).read()
    m = re.search(name + r", \[[^=]*=\s*\{(.*?)\};", txt, re.S)
    body = m.group(1)
    rows = re.findall(r"\{([^{}]*)\}", body) or [body]
    out = []
    for r in rows:
        out.append([int(v) for v in re.findall(r"ICDF\((\d+)\)", r)])
    return out


@pytest.mark.skipif(not (aomlib.available() and dav1d.available()), reason="needs libaom+dav1d binaries")
def test_anonymous_cdfs_found_in_decoder_binaries():
    aom = open(aomlib.PATH, "rb").read()
    dv = open(dav1d.PATH, "rb").read()
    for name in ("av1_default_angle_delta_cdf", "av1_default_cfl_alpha_cdf", "av1_default_cfl_sign_cdf"):
        for row in _rows(name):
            pat = struct.pack("<%dH" % len(row), *[32768 - v for v in row])
            assert dv.find(pat) >= 0 or aom.find(pat) >= 0, (name, row)


def test_smooth_weights_in_libaom():
    aom = open(aomlib.PATH, "rb").read()
    w = [255, 149, 85, 64, 255, 197, 146, 105, 73, 50, 37, 32, 255, 225, 196, 170, 145, 123, 102, 84, 68, 54, 43, 33, 26, 20, 17, 16,
         255, 240, 225, 210, 196, 182, 169, 157, 145, 133, 122, 111, 101, 92, 83, 74, 66, 59, 52, 45, 39, 34, 29, 25, 21, 17, 14, 12, 10, 9, 8, 8]
    assert aom.find(bytes(w)) >= 0
