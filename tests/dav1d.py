"""Test-only ctypes binding to the dav1d 1.5.x decoder exported by Pillow's bundled libavif.

Independent AV1 decoder used to pin every normative stage of the encoder
(SURVEY.md section 8c): emitted OBUs -> decoded planes must equal the encoder's
own reconstruction sample-exactly.  Never imported by product code.
"""
import ctypes, glob, os, sysconfig
import numpy as np

def _find():
    sp = sysconfig.get_paths()["purelib"]
    c = glob.glob(os.path.join(sp, "pillow.libs", "libavif-*.so*"))
    return c[0] if c else None

PATH = _find()
_lib = None

def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(PATH)
        _lib.dav1d_version.restype = ctypes.c_char_p
        _lib.dav1d_data_create.restype = ctypes.c_void_p
        _lib.dav1d_data_create.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    return _lib

def available():
    try:
        return PATH is not None and lib().dav1d_version() is not None
    except OSError:
        return False

INLOOP_DEBLOCK, INLOOP_CDEF, INLOOP_RESTORATION = 1, 2, 4
INLOOP_ALL = 7

def decode(obus: bytes, inloop_filters: int = INLOOP_ALL):
    """Decode one temporal unit of OBUs. Returns (planes list of np arrays, bpc, layout)."""
    L = lib()
    settings = ctypes.create_string_buffer(256)
    L.dav1d_default_settings(settings)
    ctypes.c_int.from_buffer(settings, 0).value = 1    # n_threads
    ctypes.c_int.from_buffer(settings, 4).value = 1    # max_frame_delay
    # Dav1dSettings: 6 ints, allocator (3 ptr), logger (2 ptr), strict, output_invisible, inloop_filters
    ctypes.c_int.from_buffer(settings, 72).value = inloop_filters
    ctx = ctypes.c_void_p()
    r = L.dav1d_open(ctypes.byref(ctx), settings)
    if r != 0:
        raise RuntimeError("dav1d_open failed %d" % r)
    try:
        data = ctypes.create_string_buffer(128)
        ptr = L.dav1d_data_create(data, len(obus))
        if not ptr:
            raise RuntimeError("dav1d_data_create failed")
        ctypes.memmove(ptr, obus, len(obus))
        pic = ctypes.create_string_buffer(512)
        got = False
        for _ in range(64):
            sz = ctypes.c_size_t.from_buffer(data, 8).value
            if sz > 0:
                r = L.dav1d_send_data(ctx, data)
                if r != 0 and r != -11:
                    raise RuntimeError("dav1d_send_data error %d" % r)
            r = L.dav1d_get_picture(ctx, pic)
            if r == 0:
                got = True
                break
            if r != -11:
                raise RuntimeError("dav1d_get_picture error %d" % r)
            if sz == 0:
                # drain
                r = L.dav1d_get_picture(ctx, pic)
                if r == 0:
                    got = True
                break
        if not got:
            raise RuntimeError("dav1d produced no picture")
        d = [ctypes.c_void_p.from_buffer(pic, 16 + 8 * i).value for i in range(3)]
        stride = [ctypes.c_ssize_t.from_buffer(pic, 40 + 8 * i).value for i in range(2)]
        w = ctypes.c_int.from_buffer(pic, 56).value
        h = ctypes.c_int.from_buffer(pic, 60).value
        layout = ctypes.c_int.from_buffer(pic, 64).value   # 0=I400 1=I420 2=I422 3=I444
        bpc = ctypes.c_int.from_buffer(pic, 68).value
        bps = 2 if bpc > 8 else 1
        dt = np.uint16 if bpc > 8 else np.uint8
        planes = []
        nplanes = 1 if layout == 0 else 3
        for p in range(nplanes):
            pw, ph = w, h
            if p and layout in (1, 2):
                pw = (w + 1) // 2
            if p and layout == 1:
                ph = (h + 1) // 2
            st = stride[0 if p == 0 else 1]
            buf = ctypes.string_at(d[p], st * ph)
            a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, st)
            a = a[:, : pw * bps].copy().view(dt).reshape(ph, pw)
            planes.append(a)
        L.dav1d_picture_unref(pic)
        return planes, bpc, layout
    finally:
        L.dav1d_close(ctypes.byref(ctx))
