"""Test-only access to the libaom 3.13.1 build bundled with opencv-python-headless.

The .so keeps its full .symtab, so local C reference functions (av1_inv_txfm2d_add_*_c,
av1_fwd_txfm2d_*_c, aom_*_predictor_*_c, ...) and .rodata tables can be reached by
symbol value + load base.  Used as an independent known-answer oracle for the AV1
normative arithmetic.  Never imported by product code.
"""
import ctypes, glob, os, struct, subprocess, functools, sysconfig

def _find():
    sp = sysconfig.get_paths()["purelib"]
    c = glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libaom-*.so*"))
    return c[0] if c else None

PATH = _find()

@functools.lru_cache(None)
def _symtab():
    out = subprocess.run(["nm", "-S", PATH], capture_output=True, text=True).stdout
    syms = {}
    for line in out.splitlines():
        p = line.split()
        if len(p) == 4:
            addr, size, typ, name = p
            syms.setdefault(name, (int(addr, 16), int(size, 16), typ))
        elif len(p) == 3:
            addr, typ, name = p
            syms.setdefault(name, (int(addr, 16), 0, typ))
    return syms

@functools.lru_cache(None)
def _lib():
    lib = ctypes.CDLL(PATH)
    # base = runtime address of an exported symbol - its st_value
    exp = "aom_codec_av1_dx"
    rt = ctypes.cast(getattr(lib, exp), ctypes.c_void_p).value
    base = rt - _symtab()[exp][0]
    # run-time CPU dispatch tables (function pointers in .bss) must be initialised before any
    # *_c function that calls through them (e.g. av1_round_shift_array)
    for init in ("aom_dsp_rtcd", "aom_scale_rtcd", "av1_rtcd"):
        if init in _symtab():
            ctypes.CFUNCTYPE(None)(base + _symtab()[init][0])()
    return lib, base

def available():
    return PATH is not None and "av1_inv_txfm2d_add_8x8_c" in _symtab()

def addr(name):
    lib, base = _lib()
    return base + _symtab()[name][0]

def func(name, restype, *argtypes):
    return ctypes.CFUNCTYPE(restype, *argtypes)(addr(name))

def table(name, fmt, count=None):
    """Read a .rodata/.data table as a tuple of struct-format items (e.g. 'h', 'H', 'i', 'B')."""
    a, size, _ = _symtab()[name]
    isz = struct.calcsize(fmt)
    n = count if count is not None else size // isz
    buf = ctypes.string_at(addr(name), n * isz)
    return struct.unpack("<%d%s" % (n, fmt), buf)

def raw(name):
    a, size, _ = _symtab()[name]
    return ctypes.string_at(addr(name), size)
