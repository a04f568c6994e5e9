#ifndef B200_AVIF_MUX_H_
#define B200_AVIF_MUX_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct B200AvifMuxInput {
  const uint8_t *color; size_t color_len;   // AV1 OBUs of the 4:4:4 colour image
  const uint8_t *alpha; size_t alpha_len;   // AV1 OBUs of the monochrome alpha image, or NULL/0
  const uint8_t *exif; size_t exif_len;     // optional
  uint32_t width, height;
  int depth;                                // 8 | 10
  int matrix_coefficients;                  // CICP: 0 identity(RGB), 1 BT.709, 2 unspecified, 6 BT.601, ...
  int premultiplied_alpha;
  int color_level_idx, alpha_level_idx;     // seq_level_idx_0 of the two sequence headers (av1C)
} B200AvifMuxInput;
// *out is malloc()ed. Returns 0 on success.
int b200avif_mux(const B200AvifMuxInput *in, uint8_t **out, size_t *out_len);
#ifdef __cplusplus
}
#endif
#endif
