"""The C-ABI shared library loads on a GPU-less box, exports every symbol include/b200avif.h declares,
and FAILS LOUDLY (no CPU fallback) when asked to encode without a CUDA device."""
import ctypes, os, re
import numpy as np
import pytest
import harness as H


def declared_symbols():
    txt = open(os.path.join(H.ROOT, "include", "b200avif.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200av(?:if|1)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = H.hostlib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export: " + s


def test_builder_validation_replaces_panics():
    L = H.hostlib()
    L.b200avif_encoder_new.restype = ctypes.c_void_p
    L.b200avif_encoder_with_quality.argtypes = [ctypes.c_void_p, ctypes.c_float]
    L.b200avif_encoder_with_speed.argtypes = [ctypes.c_void_p, ctypes.c_uint8]
    L.b200avif_encoder_free.argtypes = [ctypes.c_void_p]
    e = L.b200avif_encoder_new()
    assert L.b200avif_encoder_with_quality(e, 0.5) == 4 and L.b200avif_encoder_with_quality(e, 100.5) == 4
    assert L.b200avif_encoder_with_quality(e, 80.0) == 0
    assert L.b200avif_encoder_with_speed(e, 0) == 4 and L.b200avif_encoder_with_speed(e, 11) == 4 and L.b200avif_encoder_with_speed(e, 10) == 0
    L.b200avif_encoder_free(e)


def test_python_mirror_asserts_like_the_rust_builder():
    import cavif_rs_b200 as C
    with pytest.raises(AssertionError):
        C.Encoder.new().with_quality(0)
    with pytest.raises(AssertionError):
        C.Encoder.new().with_speed(11)
    with pytest.raises(AssertionError):
        C.Encoder.new().with_num_threads(0)
    e = C.Encoder.new().with_quality(70.).with_speed(4)    # the doc-test shape of ravif/src/lib.rs:1-9
    assert e is not None


def test_no_silent_cpu_fallback():
    import cavif_rs_b200 as C
    if C.device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(C.EncodingError) as ei:
        C.Encoder.new().encode_rgba(np.zeros((16, 16, 4), np.uint8))
    assert "no CPU fallback" in str(ei.value)
