"""ctypes mirror of ravif::Encoder (same names, argument meaning and error behaviour).

reference: /root/reference/ravif/src/av1encoder.rs  (builder :105-219, encode_* :243-397),
           /root/reference/ravif/src/error.rs:8-24   (Error variants)
"""
import copy
import ctypes
import enum
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcavif_b200.so")


class ColorModel(enum.IntEnum):          # av1encoder.rs:11-22
    YCbCr = 0
    RGB = 1


class AlphaColorMode(enum.IntEnum):      # av1encoder.rs:24-41
    UnassociatedDirty = 0
    UnassociatedClean = 1
    Premultiplied = 2


class BitDepth(enum.IntEnum):            # av1encoder.rs:43-50
    Auto = 0
    Eight = 8
    Ten = 10


class PixelRange(enum.IntEnum):
    Limited = 0
    Full = 1


class MatrixCoefficients(enum.IntEnum):  # rav1e enum, CICP numbering (av1encoder.rs:459-468)
    Identity = 0
    BT709 = 1
    Unspecified = 2
    BT601 = 6
    YCgCo = 8
    BT2020NCL = 9
    BT2020CL = 10


class Error(Exception):
    """ravif::Error"""


class TooFewPixels(Error):
    def __init__(self, msg="Provided buffer is smaller than width * height"):
        super().__init__(msg)


class Unsupported(Error):
    def __init__(self, msg=""):
        super().__init__("Not supported: %s" % msg)


class EncodingError(Error):
    pass


class _Image(ctypes.Structure):
    _fields_ = [("avif_file", ctypes.c_void_p), ("avif_len", ctypes.c_size_t),
                ("color_byte_size", ctypes.c_size_t), ("alpha_byte_size", ctypes.c_size_t)]


class _Input(ctypes.Structure):
    _fields_ = [("pixels", ctypes.c_void_p), ("width", ctypes.c_size_t), ("height", ctypes.c_size_t),
                ("stride_px", ctypes.c_size_t), ("channels", ctypes.c_int), ("on_device", ctypes.c_int)]


class _Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("gpu_ms_rdo", ctypes.c_double), ("gpu_ms_total", ctypes.c_double), ("host_ms_pack", ctypes.c_double),
                ("tma_launches", ctypes.c_uint64), ("tma_timeouts", ctypes.c_uint64)]


class _StageTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("ms", ctypes.c_double), ("launches", ctypes.c_uint64)]


class EncodedImage:
    """EncodedImage (av1encoder.rs:52-61)"""
    __slots__ = ("avif_file", "color_byte_size", "alpha_byte_size")

    def __init__(self, avif_file, color_byte_size, alpha_byte_size):
        self.avif_file = avif_file
        self.color_byte_size = color_byte_size
        self.alpha_byte_size = alpha_byte_size


def build(force=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(_LIB_PATH):
        r = subprocess.run(["make", "-s", "lib"], cwd=_ROOT, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libcavif_b200.so failed:\n" + r.stderr[-4000:])
    return _LIB_PATH


_lib = None


def lib():
    """Load the product library; fails loudly if it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.b200avif_encoder_new.restype = ctypes.c_void_p
        L.b200avif_encoder_clone.restype = ctypes.c_void_p
        L.b200avif_encoder_clone.argtypes = [ctypes.c_void_p]
        L.b200avif_encoder_free.argtypes = [ctypes.c_void_p]
        L.b200avif_encoder_with_quality.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.b200avif_encoder_with_alpha_quality.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.b200avif_encoder_with_quantizer.argtypes = [ctypes.c_void_p, ctypes.c_uint8]
        L.b200avif_encoder_with_alpha_quantizer.argtypes = [ctypes.c_void_p, ctypes.c_uint8]
        L.b200avif_default_num_threads.restype = ctypes.c_size_t
        L.b200avif_encoder_with_speed.argtypes = [ctypes.c_void_p, ctypes.c_uint8]
        L.b200avif_encoder_with_bit_depth.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.b200avif_encoder_with_internal_color_model.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.b200avif_encoder_with_alpha_color_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.b200avif_encoder_with_num_threads.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.b200avif_encoder_set_exif.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.b200avif_encoder_set_device.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.b200avif_encode_rgba_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Input), ctypes.c_size_t,
                                                 ctypes.POINTER(_Image), ctypes.POINTER(ctypes.c_int)]
        L.b200avif_encode_raw_planes_8_bit.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p,
                                                       ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                                       ctypes.c_int, ctypes.POINTER(_Image)]
        L.b200avif_encode_raw_planes_10_bit.argtypes = L.b200avif_encode_raw_planes_8_bit.argtypes
        L.b200avif_image_free.argtypes = [ctypes.POINTER(_Image)]
        L.b200avif_last_error_message.restype = ctypes.c_char_p
        L.b200avif_quality_to_quantizer.restype = ctypes.c_uint8
        L.b200avif_quality_to_quantizer.argtypes = [ctypes.c_float]
        L.b200avif_get_stats.argtypes = [ctypes.POINTER(_Stats)]
        L.b200avif_set_profiling.argtypes = [ctypes.c_int]
        L.b200avif_get_stage_times.restype = ctypes.c_size_t
        L.b200avif_get_stage_times.argtypes = [ctypes.POINTER(_StageTime), ctypes.c_size_t]
        L.b200avif_rate_tables_bytes.restype = ctypes.c_size_t
        L.b200avif_build_rate_tables.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.b200avif_install_rate_tables.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        _lib = L
    return _lib


def device_count():
    return lib().b200avif_device_count()


def quality_to_quantizer(q):
    return lib().b200avif_quality_to_quantizer(q)


def default_num_threads():
    """What `with_num_threads(None)` resolves to on this host (rayon::current_num_threads(), av1encoder.rs:413-415)."""
    return lib().b200avif_default_num_threads()


def shutdown():
    """Release the pooled device contexts of this process."""
    lib().b200avif_shutdown()


def set_profiling(on):
    """Per-kernel CUDA-event timing on/off (stage_times())."""
    lib().b200avif_set_profiling(1 if on else 0)


def stage_times():
    """{kernel stage: (accumulated ms between CUDA events on the launching stream, launches)} since process start."""
    buf = (_StageTime * 64)()
    n = lib().b200avif_get_stage_times(buf, 64)
    return {buf[i].name.decode(): (buf[i].ms, buf[i].launches) for i in range(min(n, 64))}


def stats():
    s = _Stats()
    lib().b200avif_get_stats(ctypes.byref(s))
    return {k: getattr(s, k) for k, _ in _Stats._fields_}


def _raise(code):
    msg = (lib().b200avif_last_error_message() or b"").decode()
    if code == 1:
        raise TooFewPixels()
    if code == 2:
        raise Unsupported(msg)
    if code == 4:
        raise ValueError(msg)
    raise EncodingError(msg or "Encoding error")


class Encoder:
    """Encoder config builder; `with_*` methods return a new configured Encoder like the Rust by-value builder."""

    def __init__(self):
        self._quality = 80.0
        self._alpha_quality = 80.0
        self._speed = 5
        self._depth = BitDepth.Auto
        self._color_model = ColorModel.YCbCr
        self._threads = None
        self._alpha_color_mode = AlphaColorMode.UnassociatedClean
        self._exif = None
        self._device = -1

    @staticmethod
    def new():
        return Encoder()

    def _copy(self):
        return copy.copy(self)

    def with_quality(self, quality):
        assert 1.0 <= quality <= 100.0      # av1encoder.rs:117 panics
        e = self._copy(); e._quality = float(quality); return e

    def with_alpha_quality(self, quality):
        assert 1.0 <= quality <= 100.0      # :146
        e = self._copy(); e._alpha_quality = float(quality); return e

    def with_speed(self, speed):
        assert 1 <= speed <= 10             # :159
        e = self._copy(); e._speed = int(speed); return e

    def with_bit_depth(self, depth):
        e = self._copy(); e._depth = BitDepth(depth); return e

    def with_depth(self, depth):            # deprecated alias, :124-127
        return self.with_bit_depth(BitDepth.Auto if depth is None else (BitDepth.Ten if depth >= 10 else BitDepth.Eight))

    def with_internal_color_model(self, color_model):
        e = self._copy(); e._color_model = ColorModel(color_model); return e

    with_internal_color_space = with_internal_color_model

    def with_num_threads(self, num_threads):
        assert num_threads is None or num_threads > 0   # :188
        e = self._copy(); e._threads = num_threads; return e

    def with_alpha_color_mode(self, mode):
        e = self._copy(); e._alpha_color_mode = AlphaColorMode(mode); return e

    def with_exif(self, exif_data):
        e = self._copy(); e._exif = bytes(exif_data); return e

    def set_exif(self, exif_data):
        self._exif = bytes(exif_data)

    def with_device(self, cuda_device):
        e = self._copy(); e._device = int(cuda_device); return e

    # ---- native handle ------------------------------------------------------------------------
    def _native(self):
        L = lib()
        h = L.b200avif_encoder_new()
        for rc in (L.b200avif_encoder_with_quality(h, self._quality),
                   L.b200avif_encoder_with_alpha_quality(h, self._alpha_quality),
                   L.b200avif_encoder_with_speed(h, self._speed),
                   L.b200avif_encoder_with_bit_depth(h, int(self._depth)),
                   L.b200avif_encoder_with_internal_color_model(h, int(self._color_model)),
                   L.b200avif_encoder_with_alpha_color_mode(h, int(self._alpha_color_mode)),
                   L.b200avif_encoder_with_num_threads(h, 0 if self._threads is None else int(self._threads)),
                   L.b200avif_encoder_set_device(h, self._device)):
            if rc:
                L.b200avif_encoder_free(h)
                _raise(rc)
        if self._exif is not None:
            L.b200avif_encoder_set_exif(h, self._exif, len(self._exif))
        return h

    @staticmethod
    def _take(img):
        data = ctypes.string_at(img.avif_file, img.avif_len) if img.avif_file else b""
        out = EncodedImage(data, img.color_byte_size, img.alpha_byte_size)
        lib().b200avif_image_free(ctypes.byref(img))
        return out

    def _describe(self, buf, channels):
        """Accept HxWxC uint8 numpy arrays (possibly row-strided views) or CUDA torch tensors."""
        if hasattr(buf, "data_ptr") and hasattr(buf, "is_cuda"):      # torch tensor
            assert buf.dim() == 3 and buf.shape[2] == channels and buf.element_size() == 1
            assert buf.stride(2) == 1 and buf.stride(1) == channels and buf.stride(0) % channels == 0
            return _Input(buf.data_ptr(), buf.shape[1], buf.shape[0], buf.stride(0) // channels, channels,
                          1 if buf.is_cuda else 0), buf
        a = np.asarray(buf)
        assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == channels
        if a.strides[2] != 1 or a.strides[1] != channels or a.strides[0] % channels:
            a = np.ascontiguousarray(a)
        return _Input(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0] // channels, channels, 0), a

    def encode_batch(self, images, channels=4):
        """Encode independent images in one device pass (the CLI's par_iter over files, src/main.rs:223).
        Returns a list of EncodedImage; raises on the first failed image."""
        L = lib()
        n = len(images)
        if n == 0:
            return []
        ins = (_Input * n)()
        keep = []
        for i, im in enumerate(images):
            ins[i], k = self._describe(im, channels)
            keep.append(k)
        outs = (_Image * n)()
        st = (ctypes.c_int * n)()
        h = self._native()
        try:
            rc = L.b200avif_encode_rgba_batch(h, ins, n, outs, st)
        finally:
            L.b200avif_encoder_free(h)
        res = [self._take(outs[i]) for i in range(n)]
        if rc:
            _raise(rc)
        return res

    def encode_rgba(self, buffer):
        """Make a new AVIF image from RGBA pixels (non-premultiplied, alpha last). av1encoder.rs:243"""
        return self.encode_batch([buffer], 4)[0]

    def encode_rgb(self, buffer):
        """Make a new AVIF image from RGB pixels. av1encoder.rs:318"""
        return self.encode_batch([buffer], 3)[0]

    def _raw(self, fn, dtype, width, height, planes, alpha, color_pixel_range, matrix_coefficients):
        L = lib()
        p = np.ascontiguousarray(np.asarray(planes, dtype=dtype).reshape(-1, 3))
        a = None if alpha is None else np.ascontiguousarray(np.asarray(alpha, dtype=dtype).reshape(-1))
        img = _Image()
        h = self._native()
        try:
            rc = fn(h, width, height, p.ctypes.data, p.shape[0], None if a is None else a.ctypes.data,
                    0 if a is None else a.shape[0], int(color_pixel_range), int(matrix_coefficients), ctypes.byref(img))
        finally:
            L.b200avif_encoder_free(h)
        if rc:
            _raise(rc)
        return self._take(img)

    def encode_raw_planes_8_bit(self, width, height, planes, alpha, color_pixel_range, matrix_coefficients):
        """planes: iterable of [u8;3]; alpha: iterable of u8 or None. av1encoder.rs:366"""
        return self._raw(lib().b200avif_encode_raw_planes_8_bit, np.uint8, width, height, planes, alpha,
                         color_pixel_range, matrix_coefficients)

    def encode_raw_planes_10_bit(self, width, height, planes, alpha, color_pixel_range, matrix_coefficients):
        """planes: iterable of [u16;3] (values 0..=1023). av1encoder.rs:390"""
        return self._raw(lib().b200avif_encode_raw_planes_10_bit, np.uint16, width, height, planes, alpha,
                         color_pixel_range, matrix_coefficients)
