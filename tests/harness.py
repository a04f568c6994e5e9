"""Shared test plumbing: ctypes views of records.h, oracle + product library loaders, synthetic images."""
import ctypes, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

class B200Frame(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "width", "height", "bit_depth", "planes", "mi_cols", "mi_rows", "sb_cols", "sb_rows",
        "tile_cols_log2", "tile_rows_log2", "tile_cols", "tile_rows")] + [
        ("tile_col_start_sb", ctypes.c_int32 * 65), ("tile_row_start_sb", ctypes.c_int32 * 65)] + [
        (n, ctypes.c_int32) for n in (
        "base_q_idx", "dc_q", "ac_q", "reduced_tx_set", "enable_cdef", "enable_restoration",
        "color_description_present", "color_primaries", "transfer_characteristics", "matrix_coefficients", "full_range",
        "partition_min", "partition_max", "rdo_tx_decision", "fine_directional_intra", "enable_cfl", "enable_intra_edge_filter",
        "lambda_q8", "lambda_satd_q8", "y_mode_mask", "uv_mode_mask", "tx_type_mask")] + [
        ("split_gate_bits", ctypes.c_int32 * 3), ("lf_search", ctypes.c_int32), ("lr_full", ctypes.c_int32), ("psy", ctypes.c_int32), ("luma_rdo_modes", ctypes.c_int32),
        ("lf_level", ctypes.c_int32 * 4), ("lf_sharpness", ctypes.c_int32),
        ("cdef_damping", ctypes.c_int32), ("cdef_bits", ctypes.c_int32),
        ("cdef_y_strength", ctypes.c_int32 * 8), ("cdef_uv_strength", ctypes.c_int32 * 8),
        ("lr_type", ctypes.c_int32 * 3)]

class B200Records(ctypes.Structure):
    _fields_ = [("mi", ctypes.c_void_p), ("eob", ctypes.c_void_p * 3), ("coef", ctypes.c_void_p * 3),
                ("stream_base", ctypes.c_void_p), ("stream_off", ctypes.c_void_p * 3),
                ("cdef_idx", ctypes.c_void_p), ("lr", ctypes.c_void_p * 3)]

RATE_TABLES_BYTES = 2 * (20 * 10 + 3 * 2 + 5 * 5 * 13 + 2 * 13 * 14 + 8 * 7 + 8 + 6 * 16 + 2 * 4 * 13 * 16 + 5 * 13 * 2 +
                         4 * 2 * 2 * 12 + 5 * 2 * 9 * 2 + 2 * 3 * 2 + 5 * 2 * 4 * 3 + 5 * 2 * 42 * 4 + 4 * 2 * 21 * 13)

DEC_DTYPE = np.dtype([("cost", "<i8"), ("y_mode", "u1"), ("tx_type", "u1"), ("uv_mode", "u1"), ("all_zero", "u1"),
                      ("angle_y", "u1"), ("angle_uv", "u1"), ("cfl_js", "u1"), ("cfl_alpha", "u1")])
LR_DTYPE = np.dtype([("use", "i1"), ("set", "i1"), ("xqd0", "i1"), ("xqd1", "i1")])
MI_DTYPE = np.dtype([("bsize", "u1"), ("y_mode", "u1"), ("uv_mode", "u1"), ("angle", "u1"),
                     ("tx_size", "u1"), ("tx_type", "u1"), ("flags", "u1"), ("cfl", "u1")])

def _run(cmd):
    """run a (make) command under an exclusive file lock: several ranks / xdist workers may import this module at once and
    two concurrent builds must never write the same shared library"""
    import fcntl
    with open(os.path.join(ROOT, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))

# libaom av1_sgr_params: (r0, r1, s0, s1) per set
SGR_PARAMS = [(2, 1, 140, 3236), (2, 1, 112, 2158), (2, 1, 93, 1618), (2, 1, 80, 1438), (2, 1, 70, 1295), (2, 1, 58, 1177),
              (2, 1, 47, 1079), (2, 1, 37, 996), (2, 1, 30, 925), (2, 1, 25, 863), (0, 1, -1, 2589), (0, 1, -1, 1618),
              (0, 1, -1, 1177), (0, 1, -1, 925), (2, 0, 56, -1), (2, 0, 22, -1)]

_oracle = None
def oracle():
    """gcc-built CPU restatement (oracle/)."""
    global _oracle
    if _oracle is None:
        if os.environ.get("B200_ORACLE_NATIVE") == "1":      # bench.py CPU arm: -O3 -march=native build made on this host
            _run(["make", "-s", "oracle-native"])
            L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_native", "liboracle_native.so"))
        else:
            _run(["make", "-s", "oracle"])
            L = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        V, I, Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        # NB: without argtypes ctypes passes Python ints as 32-bit C ints and truncates pointers.
        L.orc_encode_frame.argtypes = [V, V, V, V, V, V, V, V, V]
        L.orc_deblock_frame.argtypes = [V, V, V]
        L.orc_lf_search.argtypes = [V, V, V, V]
        L.orc_cdef_frame.argtypes = [V, V, V, V, V]
        L.orc_cdef_search_frame.argtypes = [V, V, V, V, V, V]
        L.orc_sgr_region_plane.argtypes = [V, V] + [I] * 11 + [V, V]
        L.orc_sgr_apply_region.argtypes = [V] + [I] * 11 + [V]
        L.orc_lr_frame.argtypes = [V, V, V, V, V, V]
        L.orc_convert_image.argtypes = [V, I, Z, Z, Z, I, I, V, V, V, V]
        L.orc_any_alpha.argtypes = [V, Z, Z, Z]
        L.orc_blurred_dirty_alpha.argtypes = [V, Z, Z, Z, V]
        L.orc_premultiplied_branch.argtypes = [V, Z, Z, Z, V]
        L.orc_fwd_txfm2d.argtypes = [V, I, V, I, I]
        L.orc_inv_txfm2d_add.argtypes = [V, V, I, I, I, I]
        L.orc_pixel_planes.argtypes = [ctypes.c_uint8] * 3 + [I, I, V]
        L.orc_premultiplied_minmax.argtypes = [ctypes.c_uint8, ctypes.c_uint8, V]
        L.orc_speed_tweaks_from_my_preset.argtypes = [I, I, V]
        L.orc_quality_to_quantizer.restype = ctypes.c_uint8
        L.orc_quality_to_quantizer.argtypes = [ctypes.c_float]
        L.orc_tile_target.restype = Z
        L.orc_tile_target.argtypes = [Z] * 4
        _oracle = L
    return _oracle

_host = None
def hostlib():
    """Product library. On a box without nvcc output falls back to nothing: tests that need it build it."""
    global _host
    if _host is None:
        _run(["make", "-s", "lib"])
        _host = ctypes.CDLL(os.path.join(ROOT, "cavif_rs_b200", "libcavif_b200.so"))
        _host.b200av1_frame_setup.argtypes = [ctypes.POINTER(B200Frame)] + [ctypes.c_int] * 6
        _host.b200av1_build_rate_tables.argtypes = [ctypes.c_int, ctypes.c_void_p]
        _host.b200av1_pack_frame.argtypes = [ctypes.POINTER(B200Frame), ctypes.POINTER(B200Records),
                                             ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
        V, I, Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        _host.b200av1_rdo_run.argtypes = [V] * 14
        _host.b200av1_convert_run.argtypes = [V, I, Z, Z, Z, I, I, V, V]
        _host.b200av1_seq_level_idx.argtypes = [V]
        _host.b200avif_mux.argtypes = [V, V, V]
        _host.b200av1_alpha_clean_run.argtypes = [V, Z, Z, Z, I, V]
        _host.b200_speed_tweaks.argtypes = [I, I, V]
        _host.b200avif_quality_to_quantizer.restype = ctypes.c_uint8
        _host.b200avif_quality_to_quantizer.argtypes = [ctypes.c_float]
        _host.b200avif_last_error_message.restype = ctypes.c_char_p
    return _host

def device_frame(f):
    """Copy of a frame description as a device run must receive it: closed-form deblocking levels (see oracle_filters)."""
    fd = type(f).from_buffer_copy(f)
    if hasattr(f, "_lf_base"):
        for k in range(4):
            fd.lf_level[k] = f._lf_base[k]
    return fd


def frame_setup(w, h, depth=8, planes=3, quantizer=121, tiles=1):
    f = B200Frame()
    rc = hostlib().b200av1_frame_setup(ctypes.byref(f), w, h, depth, planes, quantizer, tiles)
    assert rc == 0
    return f

def rate_tables(qidx):
    buf = np.zeros(RATE_TABLES_BYTES // 2, np.uint16)
    hostlib().b200av1_build_rate_tables(qidx, buf.ctypes.data)
    return buf

def pad_planes(planes, f):
    """Replicate-pad visible planes (h, w) to the 8-aligned coded size (mi_rows*4, mi_cols*4), as uint16."""
    H, W = f.mi_rows * 4, f.mi_cols * 4
    out = []
    for p in planes:
        a = np.asarray(p, dtype=np.uint16)
        a = np.pad(a, ((0, H - a.shape[0]), (0, W - a.shape[1])), mode="edge")
        out.append(np.ascontiguousarray(a))
    return out

def lr_units(size):
    return max((size + 32) // 64, 1)

class Records:
    def __init__(self, f):
        n_mi = f.mi_rows * f.mi_cols
        self.f = f
        self.mi = np.zeros(n_mi, MI_DTYPE)
        self.eob = [np.zeros(n_mi, np.uint16) for _ in range(3)]
        self.coef = [np.zeros(f.sb_rows * f.sb_cols * 4096, np.int16) for _ in range(3)]
        self.cdef_idx = np.zeros(f.sb_rows * f.sb_cols, np.int8)
        self.lr = [np.zeros(lr_units(f.height) * lr_units(f.width), LR_DTYPE) for _ in range(3)]
        self.rec = [np.zeros((f.mi_rows * 4, f.mi_cols * 4), np.uint16) for _ in range(3)]
        # open-loop analysis outputs per level (0: 4x4, 1: 8x8, 2: 16x16)
        self.dec = [np.zeros(max((f.mi_rows >> lv) * (f.mi_cols >> lv), 1), DEC_DTYPE) for lv in range(LEVELS)]
        self.satd = [np.zeros((max((f.mi_rows >> lv) * (f.mi_cols >> lv), 1), 14), np.uint32) for lv in range(LEVELS)]
    def c_records(self):
        r = B200Records()
        r.mi = self.mi.ctypes.data
        for p in range(3):
            r.eob[p] = self.eob[p].ctypes.data
            r.coef[p] = self.coef[p].ctypes.data
        r.cdef_idx = self.cdef_idx.ctypes.data
        for p in range(3):
            r.lr[p] = self.lr[p].ctypes.data
        return r

LEVELS = 4     # analysis levels: 4x4, 8x8, 16x16, 32x32 (records.h B200_LEVELS)

def ptrn(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])

def ptr3(arrs):
    return (ctypes.c_void_p * 3)(*[a.ctypes.data for a in arrs])

def oracle_encode(f, rt, src):
    """Run the oracle hot path on padded uint16 planes. Returns Records (with .rec)."""
    L = oracle()
    rec = Records(f)
    src3 = list(src) + [src[0]] * (3 - len(src))
    L.orc_encode_frame(ctypes.byref(f), rt.ctypes.data, ptr3(src3), ptr3(rec.rec), rec.mi.ctypes.data,
                       ptr3(rec.eob), ptr3(rec.coef), ptrn(rec.dec), ptrn(rec.satd))
    return rec

def oracle_filters(f, rec, src):
    """In-loop filter stage of the oracle: deblock the reconstruction, search the CDEF strength index per
    superblock against `src` (fills rec.cdef_idx) and apply it.  Returns (final planes, deblocked planes)."""
    L = oracle()
    deb = [r.copy() for r in rec.rec]
    src3 = list(src) + [src[0]] * (3 - len(src))
    # lf_search: the oracle replaces the closed-form levels of f by the searched ones IN PLACE (the packer needs them in the
    # frame header); the closed-form ones - what a device run of the same frame must start from - are kept in f._lf_base
    if not hasattr(f, "_lf_base"):
        f._lf_base = [int(v) for v in f.lf_level]
    for k in range(4):
        f.lf_level[k] = f._lf_base[k]
    L.orc_lf_search(ctypes.byref(f), ptr3(deb), ptr3(src3), rec.mi.ctypes.data)
    L.orc_deblock_frame(ctypes.byref(f), ptr3(deb), rec.mi.ctypes.data)
    out = [np.zeros_like(r) for r in rec.rec]
    src3 = list(src) + [src[0]] * (3 - len(src))
    L.orc_cdef_search_frame(ctypes.byref(f), ptr3(deb), ptr3(src3), ptr3(out), rec.mi.ctypes.data, rec.cdef_idx.ctypes.data)
    if any(f.lr_type[p] for p in range(f.planes)):      # loop restoration: search (fills rec.lr) + apply
        fin = [np.zeros_like(r) for r in rec.rec]
        L.orc_lr_frame(ctypes.byref(f), ptr3(src3), ptr3(deb), ptr3(out), ptr3(rec.lr), ptr3(fin))
        rec.cdef_out = out
        return fin, deb
    return out, deb


def pack(f, rec):
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    cr = rec.c_records()
    rc = hostlib().b200av1_pack_frame(ctypes.byref(f), ctypes.byref(cr), ctypes.byref(out), ctypes.byref(n))
    assert rc == 0, rc
    data = ctypes.string_at(out.value, n.value)
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    libc.free(out)
    return data

def synth_image(w, h, seed=0, alpha=False):
    """Deterministic natural-ish RGBA test frame (SURVEY.md 8d recipe: octave sinusoids + flat rectangles + noise)."""
    rng = np.random.default_rng(0xC0FFEE + seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 4), np.float32)
    lum_noise = rng.normal(0, 3, (h, w)).astype(np.float32)
    for c in range(3):
        acc = np.full((h, w), 128, np.float32)
        for o in range(4):
            fx = rng.uniform(0.5, 3) * (2 ** o) / w
            fy = rng.uniform(0.5, 3) * (2 ** o) / h
            ph = rng.uniform(0, 2 * np.pi)
            acc += (96 / 2 ** o) * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph) * 0.5
        img[..., c] = acc + lum_noise
    for _ in range(24):
        rw = int(rng.integers(16, max(17, w // 4))); rh = int(rng.integers(16, max(17, h // 4)))
        x0 = int(rng.integers(0, max(1, w - rw))); y0 = int(rng.integers(0, max(1, h - rh)))
        img[y0:y0 + rh, x0:x0 + rw, :3] = rng.integers(0, 256, 3)
    img[..., 3] = 255
    if alpha:
        cy, cx = h / 2, w / 2
        d = np.sqrt(((yy - cy) / (h * 0.45)) ** 2 + ((xx - cx) / (w * 0.45)) ** 2)
        ramp = 24.0 / min(w, h)
        img[..., 3] = np.clip((1.0 - d) / ramp, 0, 1) * 255
        for _ in range(8):
            rw = int(rng.integers(8, max(9, w // 6))); rh = int(rng.integers(8, max(9, h // 6)))
            x0 = int(rng.integers(0, max(1, w - rw))); y0 = int(rng.integers(0, max(1, h - rh)))
            img[y0:y0 + rh, x0:x0 + rw, 3] = 0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def _zunits():
    out = []
    for z in range(256):
        ur = uc = 0
        for b in range(4):
            uc |= ((z >> (2 * b)) & 1) << b
            ur |= ((z >> (2 * b + 1)) & 1) << b
        out.append((ur, uc))
    return out

_ZU = _zunits()
_BW4 = {0: 1, 3: 2, 6: 4, 9: 8}      # BLOCK_4X4, 8X8, 16X16, 32X32 -> size in 4x4 units


def expand_streams(f, mi, eob, backing, planes):
    """Compact level streams (recon.cu: per (tile superblock row, plane), appended in coding order at the row's first slot of
    the coefficient plane) -> the dense per-block layout of records.h (b200_coef_offset), for comparisons with the oracle.
    Returns (dense planes, stream lengths per plane as {(sb row, tile col): n})."""
    dense = [np.zeros_like(b) for b in backing]
    lens = [dict() for _ in range(planes)]
    mi2 = mi.reshape(f.mi_rows, f.mi_cols)
    for tr in range(f.tile_rows):
        for tc in range(f.tile_cols):
            r0, r1 = f.tile_row_start_sb[tr] * 16, min(f.tile_row_start_sb[tr + 1] * 16, f.mi_rows)
            c0, c1 = f.tile_col_start_sb[tc] * 16, min(f.tile_col_start_sb[tc + 1] * 16, f.mi_cols)
            for sr in range(r0, r1, 16):
                start = ((sr >> 4) * f.sb_cols + (c0 >> 4)) * 4096
                cur = [start] * planes
                for sc in range(c0, c1, 16):
                    for (ur, uc) in _ZU:
                        mr, mc = sr + ur, sc + uc
                        if mr >= f.mi_rows or mc >= f.mi_cols:
                            continue
                        n4 = _BW4[int(mi2[mr, mc]["bsize"])]
                        if (mr % n4) or (mc % n4):
                            continue
                        sb = (mr >> 4) * f.sb_cols + (mc >> 4)
                        zi = 0
                        for b in range(4):
                            zi |= (((mc & 15) >> b) & 1) << (2 * b)
                            zi |= (((mr & 15) >> b) & 1) << (2 * b + 1)
                        off = sb * 4096 + zi * 16
                        for p in range(planes):
                            e = int(eob[p][mr * f.mi_cols + mc])
                            dense[p][off:off + e] = backing[p][cur[p]:cur[p] + e]
                            cur[p] += e
                for p in range(planes):
                    lens[p][(sr >> 4, tc)] = cur[p] - start
    return dense, lens
