//! UNCOMPILED (no Rust toolchain in this environment): the binding a ravif maintainer would add so that
//! `ravif::Encoder::encode_rgba / encode_rgb / encode_raw_planes_*` run on the B200 path of
//! libcavif_b200.so instead of calling rav1e + avif-serialize.  1:1 with include/b200avif.h.
//!
//! Drop this file into ravif/src/, add `mod b200_ffi;` to ravif/src/lib.rs and replace the bodies of the
//! encode functions in ravif/src/av1encoder.rs as shown at the bottom.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)] pub struct b200avif_encoder { _private: [u8; 0] }

#[repr(C)]
pub struct b200avif_image {
    pub avif_file: *mut u8,
    pub avif_len: usize,
    pub color_byte_size: usize,
    pub alpha_byte_size: usize,
}

#[repr(C)]
pub struct b200avif_input {
    pub pixels: *const u8,
    pub width: usize,
    pub height: usize,
    pub stride_px: usize,
    pub channels: c_int,
    pub on_device: c_int,
}

#[link(name = "cavif_b200")]
extern "C" {
    pub fn b200avif_encoder_new() -> *mut b200avif_encoder;
    pub fn b200avif_encoder_clone(enc: *const b200avif_encoder) -> *mut b200avif_encoder;
    pub fn b200avif_encoder_free(enc: *mut b200avif_encoder);
    pub fn b200avif_encoder_with_quality(enc: *mut b200avif_encoder, quality: f32) -> c_int;
    pub fn b200avif_encoder_with_alpha_quality(enc: *mut b200avif_encoder, quality: f32) -> c_int;
    pub fn b200avif_encoder_with_quantizer(enc: *mut b200avif_encoder, quantizer: u8) -> c_int;
    pub fn b200avif_encoder_with_alpha_quantizer(enc: *mut b200avif_encoder, quantizer: u8) -> c_int;
    pub fn b200avif_encoder_with_speed(enc: *mut b200avif_encoder, speed: u8) -> c_int;
    pub fn b200avif_encoder_with_bit_depth(enc: *mut b200avif_encoder, depth: c_int) -> c_int;
    pub fn b200avif_encoder_with_internal_color_model(enc: *mut b200avif_encoder, model: c_int) -> c_int;
    pub fn b200avif_encoder_with_alpha_color_mode(enc: *mut b200avif_encoder, mode: c_int) -> c_int;
    pub fn b200avif_encoder_with_num_threads(enc: *mut b200avif_encoder, n: usize) -> c_int;
    pub fn b200avif_encoder_set_exif(enc: *mut b200avif_encoder, exif: *const u8, len: usize) -> c_int;
    pub fn b200avif_encode_rgba(enc: *const b200avif_encoder, rgba: *const u8, w: usize, h: usize, stride_px: usize,
                                out: *mut b200avif_image) -> c_int;
    pub fn b200avif_encode_rgb(enc: *const b200avif_encoder, rgb: *const u8, w: usize, h: usize, stride_px: usize,
                               out: *mut b200avif_image) -> c_int;
    pub fn b200avif_encode_raw_planes_8_bit(enc: *const b200avif_encoder, w: usize, h: usize, planes: *const u8, n_px: usize,
                                            alpha: *const u8, n_alpha: usize, range: c_int, mc: c_int,
                                            out: *mut b200avif_image) -> c_int;
    pub fn b200avif_encode_raw_planes_10_bit(enc: *const b200avif_encoder, w: usize, h: usize, planes: *const u16, n_px: usize,
                                             alpha: *const u16, n_alpha: usize, range: c_int, mc: c_int,
                                             out: *mut b200avif_image) -> c_int;
    pub fn b200avif_encode_rgba_batch(enc: *const b200avif_encoder, input: *const b200avif_input, n: usize,
                                      out: *mut b200avif_image, status: *mut c_int) -> c_int;
    pub fn b200avif_image_free(img: *mut b200avif_image);
    pub fn b200avif_last_error_message() -> *const c_char;
}

/// status code -> ravif::Error (ravif/src/error.rs:8-24)
pub fn to_error(code: c_int) -> crate::Error {
    match code {
        1 => crate::Error::TooFewPixels,
        2 => crate::Error::Unsupported("matrix coefficients"),
        _ => crate::Error::EncodingError(crate::error::EncodingErrorDetail),
    }
}

// ---- in ravif/src/av1encoder.rs ---------------------------------------------------------------------
// impl Encoder<'_> {
//     fn native(&self) -> *mut b200avif_encoder { unsafe {
//         let e = b200avif_encoder_new();
//         // the struct stores the quantizers (av1encoder.rs:67-71), already converted by with_quality (:112-120)
//         b200avif_encoder_with_quantizer(e, self.quantizer);
//         b200avif_encoder_with_alpha_quantizer(e, self.alpha_quantizer);
//         b200avif_encoder_with_speed(e, self.speed);
//         b200avif_encoder_with_bit_depth(e, match self.output_depth { BitDepth::Eight => 8, BitDepth::Ten => 10, BitDepth::Auto => 0 });
//         b200avif_encoder_with_internal_color_model(e, self.color_model as c_int);
//         b200avif_encoder_with_alpha_color_mode(e, self.alpha_color_mode as c_int);
//         b200avif_encoder_with_num_threads(e, self.threads.unwrap_or(0));
//         if let Some(x) = &self.exif { b200avif_encoder_set_exif(e, x.as_ptr(), x.len()); }
//         e } }
//
//     pub fn encode_rgba(&self, buf: Img<&[RGBA8]>) -> Result<EncodedImage, Error> { unsafe {
//         let e = self.native();
//         let mut out = std::mem::zeroed::<b200avif_image>();
//         let rc = b200avif_encode_rgba(e, buf.buf().as_ptr() as *const u8, buf.width(), buf.height(), buf.stride(), &mut out);
//         b200avif_encoder_free(e);
//         if rc != 0 { return Err(to_error(rc)); }
//         let avif_file = std::slice::from_raw_parts(out.avif_file, out.avif_len).to_vec();
//         let res = EncodedImage { avif_file, color_byte_size: out.color_byte_size, alpha_byte_size: out.alpha_byte_size };
//         b200avif_image_free(&mut out);
//         Ok(res) } }
//     // encode_rgb / encode_raw_planes_8_bit / _10_bit: identical pattern (collect the iterators into a Vec first).
// }
//
// The CLI (src/main.rs:223) keeps its `files.into_par_iter()`; for whole-directory throughput it should instead
// collect the decoded images and call b200avif_encode_rgba_batch once per GPU.
