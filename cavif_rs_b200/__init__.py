"""cavif_rs_b200 - B200-native AVIF encode path behind ravif's Encoder API.

Python mirror of `ravif::Encoder` (reference: /root/reference/ravif/src/av1encoder.rs:66-481)
over the C ABI in include/b200avif.h.  The compute happens in hand-written sm_100a CUDA
kernels inside cavif_rs_b200/libcavif_b200.so; there is no CPU fallback.
"""
from .encoder import (Encoder, EncodedImage, ColorModel, AlphaColorMode, BitDepth, PixelRange, MatrixCoefficients,
                      Error, TooFewPixels, Unsupported, EncodingError, lib, build, stats, device_count,
                      quality_to_quantizer, default_num_threads, shutdown, set_profiling, stage_times)

__all__ = ["Encoder", "EncodedImage", "ColorModel", "AlphaColorMode", "BitDepth", "PixelRange", "MatrixCoefficients",
           "Error", "TooFewPixels", "Unsupported", "EncodingError", "lib", "build", "stats", "device_count",
           "quality_to_quantizer", "default_num_threads", "shutdown", "set_profiling", "stage_times"]
