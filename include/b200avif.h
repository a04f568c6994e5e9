/* b200avif - C ABI of the B200-native AVIF encode path.
 *
 * Drop-in boundary for the encode path of kornelski/cavif-rs: every entry point below is what a
 * Rust FFI shim for `ravif::Encoder` would bind (see INTEGRATION.md and shim/ravif_ffi.rs).
 * Reference interface replaced (paths relative to /root/reference/):
 *
 *   b200avif_encoder_new / _clone / _free      Encoder::new(), #[derive(Clone)]   ravif/src/av1encoder.rs:66-110
 *   b200avif_encoder_with_quality              Encoder::with_quality              :112-120  (assert -> B200AVIF_INVALID_ARGUMENT)
 *   b200avif_encoder_with_quantizer/_alpha_quantizer   the stored fields     :67-71 (quantizer, alpha_quantizer: u8)
 *   b200avif_encoder_with_bit_depth            Encoder::with_bit_depth            :136-140
 *   b200avif_encoder_with_alpha_quality        Encoder::with_alpha_quality        :142-150
 *   b200avif_encoder_with_speed                Encoder::with_speed                :152-163
 *   b200avif_encoder_with_internal_color_model Encoder::with_internal_color_model :165-174
 *   b200avif_encoder_with_num_threads          Encoder::with_num_threads          :183-192
 *   b200avif_encoder_with_alpha_color_mode     Encoder::with_alpha_color_mode     :194-204
 *   b200avif_encoder_set_exif                  Encoder::set_exif / with_exif      :206-219
 *   b200avif_encode_rgba                       Encoder::encode_rgba               :243-275
 *   b200avif_encode_rgb                        Encoder::encode_rgb                :318-320
 *   b200avif_encode_raw_planes_8_bit/_10_bit   Encoder::encode_raw_planes_*       :366-397
 *   b200avif_encode_rgba_batch                 the file loop of the CLI           src/main.rs:223 (rayon par_iter over files)
 *   b200av1_encode_planes                      encode_to_av1 (narrowest seam)     ravif/src/av1encoder.rs:749-771
 *   b200avif_image                             EncodedImage                       :52-61
 *   status codes                               ravif::Error                       ravif/src/error.rs:8-24
 *
 * Threading: every function is re-entrant; one encoder may be shared by concurrent callers
 * (the reference's encode methods take &self).  There is NO CPU fallback: without a CUDA
 * device every encode call returns B200AVIF_ENCODING_ERROR and sets the error message.
 */
#ifndef B200AVIF_H_
#define B200AVIF_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200avif_encoder b200avif_encoder;

enum b200avif_status {
  B200AVIF_OK = 0,
  B200AVIF_TOO_FEW_PIXELS = 1,   /* Error::TooFewPixels */
  B200AVIF_UNSUPPORTED = 2,      /* Error::Unsupported(msg) */
  B200AVIF_ENCODING_ERROR = 3,   /* Error::EncodingError (also CUDA failures) */
  B200AVIF_INVALID_ARGUMENT = 4  /* replaces the builder's assert!() panics */
};
enum { B200AVIF_COLOR_YCBCR = 0, B200AVIF_COLOR_RGB = 1 };                 /* ColorModel */
enum { B200AVIF_ALPHA_UNASSOCIATED_DIRTY = 0, B200AVIF_ALPHA_UNASSOCIATED_CLEAN = 1,
       B200AVIF_ALPHA_PREMULTIPLIED = 2 };                                 /* AlphaColorMode */
enum { B200AVIF_DEPTH_AUTO = 0, B200AVIF_DEPTH_EIGHT = 8, B200AVIF_DEPTH_TEN = 10 };   /* BitDepth */
enum { B200AVIF_RANGE_LIMITED = 0, B200AVIF_RANGE_FULL = 1 };              /* rav1e PixelRange */
/* rav1e MatrixCoefficients (CICP numbering) accepted by encode_raw_planes_* (av1encoder.rs:459-468) */
enum { B200AVIF_MC_IDENTITY = 0, B200AVIF_MC_BT709 = 1, B200AVIF_MC_UNSPECIFIED = 2, B200AVIF_MC_BT601 = 6,
       B200AVIF_MC_YCGCO = 8, B200AVIF_MC_BT2020_NCL = 9, B200AVIF_MC_BT2020_CL = 10 };

/* EncodedImage: avif_file is malloc()ed by the library; release with b200avif_image_free. */
typedef struct b200avif_image {
  uint8_t *avif_file;
  size_t avif_len;
  size_t color_byte_size;
  size_t alpha_byte_size;
} b200avif_image;

/* One RGBA8 (channels = 4) or RGB8 (channels = 3) input image; stride in pixels (imgref::Img). */
typedef struct b200avif_input {
  const uint8_t *pixels;
  size_t width, height, stride_px;
  int channels;
  int on_device;   /* non-zero: `pixels` is a CUDA device pointer on the encoder's device */
} b200avif_input;

b200avif_encoder *b200avif_encoder_new(void);
b200avif_encoder *b200avif_encoder_clone(const b200avif_encoder *enc);
void b200avif_encoder_free(b200avif_encoder *enc);

int b200avif_encoder_with_quality(b200avif_encoder *enc, float quality);              /* 1..=100 */
int b200avif_encoder_with_alpha_quality(b200avif_encoder *enc, float quality);        /* 1..=100 */
/* raw struct fields (av1encoder.rs:67-71 stores quantizer / alpha_quantizer: u8; with_quality converts, :112-120) */
int b200avif_encoder_with_quantizer(b200avif_encoder *enc, uint8_t quantizer);        /* 0..=255 */
int b200avif_encoder_with_alpha_quantizer(b200avif_encoder *enc, uint8_t quantizer);  /* 0..=255 */
int b200avif_encoder_with_speed(b200avif_encoder *enc, uint8_t speed);                /* 1..=10 */
int b200avif_encoder_with_bit_depth(b200avif_encoder *enc, int depth);                /* 0 auto(=10) | 8 | 10 */
int b200avif_encoder_with_internal_color_model(b200avif_encoder *enc, int color_model);
int b200avif_encoder_with_alpha_color_mode(b200avif_encoder *enc, int mode);
/* 0 = None = rayon::current_num_threads() of the host (av1encoder.rs:413-415,666): RAYON_NUM_THREADS, else the
 * usable cores.  It only sets the AV1 tile count, min(threads, area / min_tile_size^2), so like the reference the
 * default output depends on the host; pass n >= the cap for the most tiles (most device parallelism). */
int b200avif_encoder_with_num_threads(b200avif_encoder *enc, size_t num_threads);
size_t b200avif_default_num_threads(void);   /* what 0 / None resolves to on this host */
int b200avif_encoder_set_exif(b200avif_encoder *enc, const uint8_t *exif, size_t len);/* copied */
int b200avif_encoder_set_device(b200avif_encoder *enc, int cuda_device);              /* default: current device */

int b200avif_encode_rgba(const b200avif_encoder *enc, const uint8_t *rgba, size_t width, size_t height,
                         size_t stride_px, b200avif_image *out);
int b200avif_encode_rgb(const b200avif_encoder *enc, const uint8_t *rgb, size_t width, size_t height,
                        size_t stride_px, b200avif_image *out);
/* planes: n_px interleaved [P;3] triples, alpha: n_alpha samples or NULL (iterator semantics:
 * fewer than width*height items -> B200AVIF_TOO_FEW_PIXELS). */
int b200avif_encode_raw_planes_8_bit(const b200avif_encoder *enc, size_t width, size_t height, const uint8_t *planes,
                                     size_t n_px, const uint8_t *alpha, size_t n_alpha, int color_pixel_range,
                                     int matrix_coefficients, b200avif_image *out);
int b200avif_encode_raw_planes_10_bit(const b200avif_encoder *enc, size_t width, size_t height, const uint16_t *planes,
                                      size_t n_px, const uint16_t *alpha, size_t n_alpha, int color_pixel_range,
                                      int matrix_coefficients, b200avif_image *out);

/* Batch of independent images on ONE GPU (the caller shards batches over GPUs, one process per
 * GPU).  status[i] receives the per-image result; the return value is the first non-OK status
 * or B200AVIF_OK.  A failed image does not poison the rest of the batch. */
int b200avif_encode_rgba_batch(const b200avif_encoder *enc, const b200avif_input *in, size_t n, b200avif_image *out,
                               int *status);

void b200avif_image_free(b200avif_image *img);
const char *b200avif_last_error_message(void);   /* thread-local */

/* ravif::quality_to_quantizer (av1encoder.rs:526-530), exposed for callers/tests. */
uint8_t b200avif_quality_to_quantizer(float quality);

/* ---- narrowest seam: encode_to_av1 (av1encoder.rs:749-771) ------------------------------------ */
typedef struct b200av1_config {
  size_t width, height;
  int bit_depth;        /* 8 | 10 */
  int quantizer;        /* 0..255 */
  int speed;            /* 1..10, expanded with SpeedTweaks::from_my_preset (av1encoder.rs:555-605) */
  size_t threads;       /* 0 = None */
  int pixel_range;      /* B200AVIF_RANGE_* */
  int monochrome;       /* ChromaSampling::Cs400 (alpha) vs Cs444 */
  int has_color_description, matrix_coefficients;   /* ColorDescription{SRGB, BT709, mc} (:407-411) */
} b200av1_config;
/* planes: host pointers to uint16 samples (1 or 3 planes), strides in samples. *obu is malloc()ed. */
int b200av1_encode_planes(const b200av1_config *cfg, const uint16_t *const planes[3], const size_t strides[3],
                          uint8_t **obu, size_t *obu_len);

/* ---- introspection used by the benchmark and the parity tests ------------------------------------ */
typedef struct b200avif_stats {
  uint64_t kernel_launches;     /* CUDA kernels launched by this library since process start */
  uint64_t h2d_bytes, d2h_bytes;
  double gpu_ms_rdo;            /* accumulated CUDA-event time of the RDO kernel */
  double gpu_ms_total;          /* accumulated CUDA-event time of all device work */
  double host_ms_pack;          /* accumulated wall time of symbol packing + muxing */
  uint64_t tma_launches;        /* fused filter launches whose tiles were staged by TMA (cp.async.bulk.tensor) */
  uint64_t tma_timeouts;        /* TMA tiles that did not land in time and were re-staged with ordinary loads (0 = healthy) */
} b200avif_stats;
void b200avif_get_stats(b200avif_stats *out);
/* Per-kernel CUDA-event times (accumulated since process start while profiling is on): the time between consecutive
 * events on the launching stream, i.e. the duration of the kernel launched in between. */
typedef struct b200avif_stage_time { char name[48]; double ms; uint64_t launches; } b200avif_stage_time;
void b200avif_set_profiling(int on);
size_t b200avif_get_stage_times(b200avif_stage_time *out, size_t cap);   /* returns the number of stages */
/* Multi-GPU batches (one process per GPU): the quantiser / rate tables of a batch are built ONCE (rank 0,
 * b200avif_build_rate_tables into rate_tables_bytes() of host memory), broadcast over NCCL, and installed by every rank
 * from the DEVICE buffer it received; encodes on that device with that base_q_idx then use the installed copy instead of
 * building their own.  (BASELINE.json north_star: "NCCL only to broadcast quantizer tables".) */
size_t b200avif_rate_tables_bytes(void);
int b200avif_build_rate_tables(int base_q_idx, void *host_out);
int b200avif_install_rate_tables(int cuda_device, int base_q_idx, const void *device_tables, size_t bytes);
/* Releases the pooled device contexts (streams, arenas, cached rate tables) of this process. */
void b200avif_shutdown(void);
int b200avif_device_count(void);   /* 0 when no CUDA device is usable */

#ifdef __cplusplus
}
#endif
#endif /* B200AVIF_H_ */
