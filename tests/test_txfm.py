"""1-D / 2-D AV1 integer transforms of the oracle pinned bit-exactly to libaom 3.13.1's C reference
functions (called by symbol address in the unstripped .so bundled with opencv-python-headless)."""
import ctypes
import numpy as np
import pytest
import aomlib
import harness as H

pytestmark = pytest.mark.skipif(not aomlib.available(), reason="libaom with symbols not found")
P16 = ctypes.POINTER(ctypes.c_int16); P32 = ctypes.POINTER(ctypes.c_int32); PU16 = ctypes.POINTER(ctypes.c_uint16)


@pytest.mark.parametrize("n", [4, 8, 16, 32])
def test_2d_against_libaom(n):
    L = H.oracle()
    rng = np.random.default_rng(n)
    fw = aomlib.func('av1_fwd_txfm2d_%dx%d_c' % (n, n), None, P16, P32, ctypes.c_int, ctypes.c_int, ctypes.c_int)
    iv = aomlib.func('av1_inv_txfm2d_add_%dx%d_c' % (n, n), None, P32, PU16, ctypes.c_int, ctypes.c_int, ctypes.c_int)
    for tt in ((0, 1, 2, 3, 9) if n < 32 else (0, 9)):
        for bd in (8, 10):
            amp = (1 << bd) - 1
            for t in range(25):
                x = rng.integers(-amp, amp + 1, (n, n)).astype(np.int16)
                if t % 3 == 0:
                    x = (x // 8).astype(np.int16)
                o1 = np.zeros((n, n), np.int32); o2 = np.zeros((n, n), np.int32)
                fw(x.ctypes.data_as(P16), o1.ctypes.data_as(P32), n, tt, bd)
                L.orc_fwd_txfm2d(x.ctypes.data_as(P16), n, o2.ctypes.data_as(P32), n, tt)
                # libaom >= 3.6 stores coefficients column-major: compare with the transpose
                assert np.array_equal(o1.T, o2), (n, tt, bd)
                q = (o2 // 16 * 16).astype(np.int32)
                pred = rng.integers(0, amp + 1, (n, n)).astype(np.uint16)
                d2 = pred.copy(); L.orc_inv_txfm2d_add(q.ctypes.data_as(P32), d2.ctypes.data_as(PU16), n, n, tt, bd)
                qt = np.ascontiguousarray(q.T)
                d1 = pred.copy(); iv(qt.ctypes.data_as(P32), d1.ctypes.data_as(PU16), n, tt, bd)
                assert np.array_equal(d1, d2), (n, tt, bd)


def test_scan_tables_are_transposed_libaom():
    import re, os
    txt = open(os.path.join(H.ROOT, "cavif_rs_b200/csrc/av1_tables.h")).read()
    for n in (4, 8, 16, 32):
        m = re.search(r"av1_default_scan_%dx%d, \[\d+\]\) = \{([^}]*)\}" % (n, n), txt)
        spec = [int(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()]
        aom = aomlib.table("default_scan_%dx%d" % (n, n), "h")
        assert spec == [(p % n) * n + (p // n) for p in aom]
        assert sorted(spec) == list(range(n * n))
