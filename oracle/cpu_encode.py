"""ORACLE DRIVER (test infrastructure): whole-image CPU encode = oracle hot path (oracle/*.c: colour
conversion, alpha clean-up, RDO, reconstruction) + the product's host-side symbol packer and muxer.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
It exists to (a) check the device path end to end and (b) time the same algorithm on the host cores.
"""
import ctypes
import os
import sys
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(_ROOT, "tests"))
import harness as H   # noqa: E402


class SpeedTweaks(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("speed_preset", "fast_deblock", "reduced_tx_set", "tx_domain_distortion", "tx_domain_rate",
                                            "encode_bottomup", "rdo_tx_decision", "cdef", "lrf", "sgr_complexity_full", "use_satd_subpel",
                                            "inter_tx_split", "fine_directional_intra", "complex_prediction_modes", "partition_min",
                                            "partition_max", "min_tile_size")]


def _lib():
    L = H.oracle()
    L.orc_quality_to_quantizer.restype = ctypes.c_uint8
    L.orc_quality_to_quantizer.argtypes = [ctypes.c_float]
    L.orc_tile_target.restype = ctypes.c_size_t
    L.orc_tile_target.argtypes = [ctypes.c_size_t] * 4
    L.orc_blurred_dirty_alpha.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    L.orc_convert_image.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                    ctypes.c_int] + [ctypes.c_void_p] * 4
    L.orc_any_alpha.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t]
    return L


def configure(w, h, depth, planes, quantizer, speed, threads=None):
    """Frame description exactly as the product's configure_frame() builds it (api.cpp)."""
    L = _lib()
    tw = SpeedTweaks()
    L.orc_speed_tweaks_from_my_preset(speed, quantizer, ctypes.byref(tw))
    cap = (w * h) // (tw.min_tile_size ** 2)
    if threads is None:      # ravif `None` = rayon::current_num_threads() of this host (av1encoder.rs:413-415, :666)
        hl = H.hostlib()
        hl.b200avif_default_num_threads.restype = ctypes.c_size_t
        threads = hl.b200avif_default_num_threads()
    tiles = min(threads, cap)
    tiles = min(max(tiles, 1), 4096)
    f = H.frame_setup(w, h, depth, planes, quantizer, tiles)
    f.reduced_tx_set = tw.reduced_tx_set
    f.rdo_tx_decision = tw.rdo_tx_decision
    f.partition_max = min(tw.partition_max, 32)
    f.partition_min = min(tw.partition_min, f.partition_max)
    f.fine_directional_intra = tw.fine_directional_intra
    f.enable_intra_edge_filter = 1
    f.split_gate_bits[0] = 0 if speed <= 2 else 60
    f.split_gate_bits[1] = 0 if speed <= 2 else 120
    f.split_gate_bits[2] = 0 if speed <= 2 else 200
    f.psy = 1          # Tune::Psychovisual, always (av1encoder.rs:694)
    f.lf_search = 0 if tw.fast_deblock else 1
    f.lr_full = 1 if tw.sgr_complexity_full else 0
    f.luma_rdo_modes = 4 if tw.complex_prediction_modes else 2
    f.enable_cdef = 1 if tw.cdef else 0
    f.enable_restoration = 1 if tw.lrf else 0
    for p in range(planes):
        f.lr_type[p] = 1 if tw.lrf else 0
    return f


def convert(rgba, depth, color_model=0):
    """uint8 HxWx{3,4} -> ([y,u,v] uint16 HxW, alpha or None, any_alpha) with the oracle's colour arithmetic."""
    L = _lib()
    a = np.ascontiguousarray(rgba)
    h, w, ch = a.shape
    pl = [np.zeros((h, w), np.uint16) for _ in range(4)]
    L.orc_convert_image(a.ctypes.data, ch, w, h, w, color_model, depth, pl[0].ctypes.data, pl[1].ctypes.data, pl[2].ctypes.data,
                        pl[3].ctypes.data if ch == 4 else None)
    anya = bool(L.orc_any_alpha(a.ctypes.data, w, h, w)) if ch == 4 else False
    return pl[:3], (pl[3] if ch == 4 else None), anya


def encode_frame(planes, w, h, depth, quantizer, speed, threads=None, color=True, mc=6):
    """One AV1 frame on the CPU: returns (obu bytes, Records incl. reconstruction, frame)."""
    f = configure(w, h, depth, len(planes), quantizer, speed, threads)
    f.full_range = 1
    if color:
        f.color_description_present = 1; f.color_primaries = 1; f.transfer_characteristics = 13; f.matrix_coefficients = mc
    src = H.pad_planes(planes, f)
    rec = H.oracle_encode(f, H.rate_tables(f.base_q_idx), src)
    rec.final, rec.deblocked = H.oracle_filters(f, rec, src)      # also fills rec.cdef_idx
    return H.pack(f, rec), rec, f


def encode_rgba(rgba, quality=80.0, speed=4, depth=10, alpha_quality=80.0, color_model=0, alpha_mode=1, threads=None):
    """Whole ravif::Encoder::encode_rgba flow on the CPU (av1encoder.rs:243-275). Returns dict."""
    L = _lib()
    a = np.ascontiguousarray(rgba)
    h, w, ch = a.shape
    if ch == 4 and alpha_mode == 1:
        out = np.zeros_like(a)
        if L.orc_blurred_dirty_alpha(a.ctypes.data, w, h, w, out.ctypes.data):
            a = out
    elif ch == 4 and alpha_mode == 2:      # AlphaColorMode::Premultiplied (av1encoder.rs:277-299)
        out = np.zeros_like(a)
        L.orc_premultiplied_branch(a.ctypes.data, w, h, w, out.ctypes.data)
        a = out
    q = L.orc_quality_to_quantizer(quality)
    aq = L.orc_quality_to_quantizer(alpha_quality)
    planes, alpha, anya = convert(a, depth, color_model)
    mc = 6 if color_model == 0 else 0
    col, crec, cf = encode_frame(planes, w, h, depth, q, speed, threads, True, mc)
    res = {"color_obu": col, "color_rec": crec, "color_frame": cf, "alpha_obu": None}
    if anya:
        al, arec, af = encode_frame([alpha], w, h, depth, aq, speed, threads, False)
        res.update(alpha_obu=al, alpha_rec=arec, alpha_frame=af)
    return res
