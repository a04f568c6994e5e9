"""Kernel logic check without a GPU: rdo.cu compiled as host C++ on the test-only CUDA emulator
(tests/emu/cuda_emu.h: one fibre per CUDA thread) must reproduce the oracle's records and
reconstruction bit-exactly.  The same comparison runs against the real kernel in test_gpu_parity.py."""
import ctypes, os, subprocess
import numpy as np
import pytest
import harness as H


class RdoJob(ctypes.Structure):
    _fields_ = [("frame", H.B200Frame), ("rate", ctypes.c_void_p), ("stride", ctypes.c_int32), ("reserved_", ctypes.c_int32),
                ("src", ctypes.c_void_p * 3), ("rec", ctypes.c_void_p * 3),
                ("mi", ctypes.c_void_p), ("eob", ctypes.c_void_p * 3), ("coef", ctypes.c_void_p * 3), ("stream_len", ctypes.c_void_p * 3),
                ("dec", ctypes.c_void_p * H.LEVELS), ("satd", ctypes.c_void_p * H.LEVELS), ("list", ctypes.c_void_p * H.LEVELS), ("list_n", ctypes.c_void_p),
                ("cdef_idx", ctypes.c_void_p), ("fin", ctypes.c_void_p * 3), ("lf_sse", ctypes.c_void_p), ("lf_out", ctypes.c_void_p), ("dbrows", ctypes.c_void_p * 3), ("deb_full", ctypes.c_void_p * 3),
                ("lr", ctypes.c_void_p * 3), ("lr_out", ctypes.c_void_p * 3)]


class TileRef(ctypes.Structure):
    _fields_ = [("job", ctypes.c_int32), ("tile_row", ctypes.c_int16), ("tile_col", ctypes.c_int16)]


def emu_lib():
    out = os.path.join(H.ROOT, "build", "libemu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = [os.path.join(H.ROOT, p) for p in ("tests/emu/emu_main.cpp", "tests/emu/cuda_emu.h", "cavif_rs_b200/csrc/rdo.cu", "cavif_rs_b200/csrc/recon.cu", "cavif_rs_b200/csrc/av1_dev.cuh",
                                             "cavif_rs_b200/csrc/filters.cu", "cavif_rs_b200/csrc/rdo.h", "cavif_rs_b200/csrc/records.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in src):
        r = subprocess.run(["g++", "-O1", "-std=c++17", "-DB200_EMU", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-Itests/emu",
                            "-Icavif_rs_b200/csrc", "-o", out, "tests/emu/emu_main.cpp"], cwd=H.ROOT, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    return ctypes.CDLL(out)


def compare(ref, out, planes):
    for p in range(planes):
        assert np.array_equal(ref.rec[p], out.rec[p]), "reconstruction plane %d" % p
        assert np.array_equal(ref.eob[p], out.eob[p]), "eob plane %d" % p
        assert np.array_equal(ref.coef[p], out.coef[p]), "coefficients plane %d" % p
    assert np.array_equal(ref.mi, out.mi), "mode info"


def compare_analysis(ref, out, f):
    """stage A outputs: SATD tables and per-block decisions of every analysed level"""
    for lv in range(H.LEVELS):
        n = 4 << lv
        if n < min(f.partition_min, 8) or n > f.partition_max:
            continue
        assert np.array_equal(ref.satd[lv], out.satd[lv]), "SATD table level %d" % lv
        for k in ("cost", "y_mode", "tx_type", "uv_mode", "all_zero", "angle_y", "angle_uv", "cfl_js", "cfl_alpha"):
            assert np.array_equal(ref.dec[lv][k], out.dec[lv][k]), "decision %s level %d" % (k, lv)


@pytest.mark.parametrize("rows", [1, 2, 3, 4])
@pytest.mark.parametrize("w,h,depth,planes,tiles,kw", [
    (64, 48, 8, 3, 1, {}),
    (40, 72, 10, 1, 2, dict(partition_min=8, rdo_tx_decision=0)),
    (72, 40, 8, 3, 1, dict(partition_min=16, partition_max=16, rdo_tx_decision=0)),   # forced 8x8 splits at the edges
    (200, 136, 8, 3, 1, dict(partition_min=8)),
    (96, 80, 8, 3, 1, dict(split_gates=(44, 96))),       # early termination gates of the fast presets
    (161, 100, 8, 3, 1, dict(lr=1, partition_min=8)),    # loop restoration: odd sizes, merged last unit row, two stripes
    (95, 70, 10, 1, 1, dict(lr=1, partition_min=16)),    # one 95-wide unit, 10 bit, monochrome
    (136, 104, 8, 3, 1, dict(partition_max=32)),         # 32x32 level on top (partition range of speed <= 2)
    (96, 64, 10, 3, 1, dict(partition_max=32, partition_min=8, lr=1)),
    (120, 88, 10, 1, 2, dict(split_gates=(60, 120), partition_min=8, lr=1)),   # monochrome (alpha) frame of the fast presets
    (136, 104, 8, 3, 1, dict(enable_intra_edge_filter=1)),                     # intra edge filter / upsampling
    (96, 96, 10, 3, 1, dict(enable_intra_edge_filter=1, partition_max=32, y_mode_mask=0x1fe)),   # directional modes only, up to 32x32
    # coded size 8 mod 16 and wide: the root-region count (129) exceeds n_mi / 16 + 64 (128); every region must be decided
    (2056, 8, 8, 1, 1, dict(partition_min=8, rdo_tx_decision=0)),
    (136, 88, 8, 3, 1, dict(reduced_tx_set=0)),
    (200, 136, 8, 3, 1, dict(psy=1)),                               # Tune::Psychovisual: variance-aware luma error per 8x8
    (136, 104, 10, 3, 1, dict(psy=1, partition_max=32, reduced_tx_set=0, lr=1)),
    (200, 136, 8, 3, 1, dict(lf_search=1, lr=1)),                   # deblocking level search + restoration
    (136, 104, 10, 1, 1, dict(lf_search=1, lr=1, lr_full=1, partition_min=8)),
    (136, 104, 8, 3, 1, dict(luma_rdo_modes=4, psy=1, reduced_tx_set=0)),          # four fully evaluated luma modes (complex_prediction_modes)
    (96, 64, 10, 3, 2, dict(luma_rdo_modes=4, partition_max=32, partition_min=8)),   # all 16 self-guided sets, monochrome                    # 7-type transform set on 4x4 / 8x8 (V_DCT, H_DCT: 1-D classes)
    (96, 72, 10, 3, 2, dict(reduced_tx_set=0, partition_max=32, tx_type_mask=0x61)),   # DCT_DCT + the 1-D DCTs where they exist
      # 4 x 3 superblocks: the reconstruction wavefront waits on its neighbours
])
def test_emulated_kernel_matches_oracle(w, h, depth, planes, tiles, kw, rows):
    E = emu_lib()
    f = H.frame_setup(w, h, depth, planes, 121, tiles)
    for k, v in kw.items():
        if k == "split_gates":
            f.split_gate_bits[0], f.split_gate_bits[1] = v
        elif k == "lr":
            f.enable_restoration = 1
            for p in range(planes):
                f.lr_type[p] = 1
        else:
            setattr(f, k, v)
    rt = H.rate_tables(f.base_q_idx)
    img = H.synth_image(w, h, 3)
    src = H.pad_planes([img[..., c].astype(np.uint16) * (4 if depth == 10 else 1) for c in range(planes)], f)
    ref = H.oracle_encode(f, rt, src)
    out = H.Records(f)
    lf_sse = np.zeros(12, np.uint64); lf_out = np.zeros(4, np.int32)
    job = RdoJob(); job.frame = f; job.rate = rt.ctypes.data
    job.lf_sse = lf_sse.ctypes.data; job.lf_out = lf_out.ctypes.data
    # the kernels work on planes of the native sample width: uint8 for 8-bit video, uint16 for 10-bit
    pix = np.uint8 if depth == 8 else np.uint16
    job.stride = f.mi_cols * 4
    src3 = [np.ascontiguousarray(a.astype(pix)) for a in (list(src) + [src[0]] * (3 - len(src)))]
    out.rec = [np.zeros(r.shape, pix) for r in out.rec]
    for p in range(3):
        job.src[p] = src3[p].ctypes.data; job.rec[p] = out.rec[p].ctypes.data
        job.eob[p] = out.eob[p].ctypes.data; job.coef[p] = out.coef[p].ctypes.data
    slen = [np.zeros(f.sb_rows * f.tile_cols, np.uint32) for _ in range(3)]
    for p in range(3):
        job.stream_len[p] = slen[p].ctypes.data
    job.mi = out.mi.ctypes.data
    lists = [np.zeros(max((f.mi_rows >> lv) * (f.mi_cols >> lv), 1), np.uint32) for lv in range(H.LEVELS)]
    list_n = np.zeros(H.LEVELS, np.uint32)
    job.list_n = list_n.ctypes.data
    for lv in range(H.LEVELS):
        job.list[lv] = lists[lv].ctypes.data
        job.dec[lv] = out.dec[lv].ctypes.data; job.satd[lv] = out.satd[lv].ctypes.data
    nt = f.tile_cols * f.tile_rows
    trs = (TileRef * nt)()
    for i in range(nt):
        trs[i].job = 0; trs[i].tile_row = i // f.tile_cols; trs[i].tile_col = i % f.tile_cols
    fin = [np.zeros_like(r) for r in out.rec]
    for p in range(3):
        job.fin[p] = fin[p].ctypes.data
    job.cdef_idx = out.cdef_idx.ctypes.data
    lr_out = [np.zeros_like(r) for r in out.rec]
    for p in range(3):
        job.lr[p] = out.lr[p].ctypes.data; job.lr_out[p] = lr_out[p].ctypes.data
    E.b200_rdo_launch_emu.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    E.b200_filters_launch_emu.argtypes = [ctypes.c_void_p, ctypes.c_int]
    E.b200_rdo_launch_emu(ctypes.byref(job), trs, nt, rows)
    compare_analysis(ref, out, f)
    # the reconstruction chain emits compact level streams: expand them to the dense per-block layout of the oracle
    out.coef, lens = H.expand_streams(f, out.mi, out.eob, out.coef, planes)
    for p in range(planes):
        for (sbr, tc), n in lens[p].items():
            assert slen[p][sbr * f.tile_cols + tc] == n, "stream length plane %d region %d,%d" % (p, sbr, tc)
    compare(ref, out, planes)
    # in-loop filter stage: deblock in place, CDEF strength search, CDEF output
    ref_fin, ref_deb = H.oracle_filters(f, ref, src)      # lf_search: f.lf_level now holds the searched levels, job.frame the closed-form ones
    deb = [np.zeros_like(r) for r in out.rec]
    dbrows = [np.zeros((f.sb_rows * 4, f.mi_cols * 4), pix) for _ in range(3)]
    for p in range(3):
        job.deb_full[p] = deb[p].ctypes.data; job.dbrows[p] = dbrows[p].ctypes.data
    E.b200_filters_launch_emu(ctypes.byref(job), f.sb_rows * f.sb_cols)
    if f.lf_search:
        assert [int(v) for v in lf_out] == [int(v) for v in f.lf_level], "deblocking levels"
    uses_lr = any(f.lr_type[p] for p in range(planes))
    ref_cdef = ref.cdef_out if uses_lr else ref_fin
    for p in range(planes):
        assert np.array_equal(deb[p], ref_deb[p]), "deblocked plane %d" % p
        assert np.array_equal(fin[p], ref_cdef[p]), "CDEF output plane %d" % p
    assert np.array_equal(out.cdef_idx, ref.cdef_idx), "cdef index"
    if uses_lr:
        for p in range(planes):
            assert np.array_equal(out.lr[p], ref.lr[p]), "restoration units plane %d" % p
            assert np.array_equal(lr_out[p][:h, :w], ref_fin[p][:h, :w]), "restored plane %d" % p
