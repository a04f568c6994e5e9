/* ORACLE (test infrastructure, never linked into the product path).
 *
 * Plain-C restatement of the arithmetic that lives IN the reference tree
 * (/root/reference/ravif/src/av1encoder.rs and dirtyalpha.rs).  Each function cites the
 * reference lines it follows.  Pinned by: exhaustive 2^24 colour sweep facts recorded in
 * SURVEY.md App. A.3, the quality->quantizer table of App. A.1 and the `preminmax`
 * known-answer test (dirtyalpha.rs:126-135).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).  fmaf()/roundf() are
 * used exactly where the Rust source has mul_add()/round().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* av1encoder.rs:501 */
static const float BT601[3] = {0.2990f, 0.5870f, 0.1140f};

/* av1encoder.rs:485-488 */
static inline uint16_t to_ten(uint8_t x) { return (uint16_t)(((uint16_t)x << 2) | ((uint16_t)x >> 6)); }
uint16_t orc_to_ten(uint8_t x) { return to_ten(x); }

/* av1encoder.rs:504-512: f32 BT.601 full-range with explicit FMAs and round-half-away. */
void orc_rgb_to_ycbcr_f32(uint8_t r, uint8_t g, uint8_t b, int depth, float out[3]) {
  const float max_value = (float)((1 << depth) - 1);
  const float scale = max_value / 255.f;
  const float shift = roundf(max_value * 0.5f);
  const float y = fmaf(scale * BT601[2], (float)b, fmaf(scale * BT601[0], (float)r, scale * BT601[1] * (float)g));
  const float cb = fmaf(fmaf((float)b, scale, -y), 0.5f / (1.f - BT601[2]), shift);
  const float cr = fmaf(fmaf((float)r, scale, -y), 0.5f / (1.f - BT601[0]), shift);
  out[0] = roundf(y); out[1] = roundf(cb); out[2] = roundf(cr);
}

/* Rust `as u8` / `as u16` from f32: saturating (av1encoder.rs:514-524). */
static inline uint8_t sat_u8(float v) { return v <= 0.f ? 0 : (v >= 255.f ? 255 : (uint8_t)v); }
static inline uint16_t sat_u16(float v) { return v <= 0.f ? 0 : (v >= 65535.f ? 65535 : (uint16_t)v); }

/* One pixel -> three samples.  color_model: 0 = YCbCr (BT601), 1 = RGB (planes G,B,R). depth 8|10.
 * av1encoder.rs:485-498 (gbr), :514-524 (ycbcr). */
void orc_pixel_planes(uint8_t r, uint8_t g, uint8_t b, int color_model, int depth, uint16_t out[3]) {
  if (color_model == 1) {
    if (depth == 8) { out[0] = g; out[1] = b; out[2] = r; }
    else { out[0] = to_ten(g); out[1] = to_ten(b); out[2] = to_ten(r); }
    return;
  }
  float f[3];
  orc_rgb_to_ycbcr_f32(r, g, b, depth, f);
  if (depth == 8) { out[0] = sat_u8(f[0]); out[1] = sat_u8(f[1]); out[2] = sat_u8(f[2]); }
  else { out[0] = sat_u16(f[0]); out[1] = sat_u16(f[1]); out[2] = sat_u16(f[2]); }
}

/* Whole image: interleaved RGBA8/RGB8 (stride in pixels) -> planar samples, row-major w*h each,
 * stored as u16 regardless of depth (tests compare values).  Follows the lazy iterators of
 * av1encoder.rs:259-271 (rgba) / :332-347 (rgb) drained row-major by init_frame_3/_1 (:710-746).
 * alpha_out may be NULL.  Alpha is px.a (8-bit) or to_ten(px.a) (10-bit), :262,:271. */
void orc_convert_image(const uint8_t *px, int channels, size_t w, size_t h, size_t stride_px, int color_model,
                       int depth, uint16_t *p0, uint16_t *p1, uint16_t *p2, uint16_t *alpha_out) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const uint8_t *p = px + (y * stride_px + x) * channels;
      uint16_t o[3];
      orc_pixel_planes(p[0], p[1], p[2], color_model, depth, o);
      p0[y * w + x] = o[0]; p1[y * w + x] = o[1]; p2[y * w + x] = o[2];
      if (alpha_out && channels == 4) alpha_out[y * w + x] = depth == 8 ? p[3] : to_ten(p[3]);
    }
}

/* av1encoder.rs:246: use_alpha = any(a != 255) */
int orc_any_alpha(const uint8_t *rgba, size_t w, size_t h, size_t stride_px) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++)
      if (rgba[(y * stride_px + x) * 4 + 3] != 255) return 1;
  return 0;
}

/* av1encoder.rs:526-530 */
uint8_t orc_quality_to_quantizer(float quality) {
  const float q = quality / 100.f;
  float x;
  if (q >= 0.82f) x = (1.f - q) * 2.6f;
  else if (q > 0.25f) x = fmaf(q, -0.5f, 1.f - 0.125f);
  else x = 1.f - q;
  return sat_u8(roundf(x * 255.f));
}

/* av1encoder.rs:533-552 / :555-605.  Option<bool> fields are always Some() in from_my_preset
 * except tx_domain_distortion (None): encoded here as -1. */
typedef struct {
  int speed_preset;
  int fast_deblock, reduced_tx_set, tx_domain_distortion, tx_domain_rate, encode_bottomup, rdo_tx_decision;
  int cdef, lrf, sgr_complexity_full, use_satd_subpel, inter_tx_split, fine_directional_intra;
  int complex_prediction_modes;
  int partition_min, partition_max;
  int min_tile_size;
} orc_speed_tweaks;

void orc_speed_tweaks_from_my_preset(int speed, int quantizer, orc_speed_tweaks *t) {
  const int low_quality = quantizer < orc_quality_to_quantizer(55.f);
  const int high_quality = quantizer > orc_quality_to_quantizer(80.f);
  const int max_block_size = high_quality ? 16 : 64;
  memset(t, 0, sizeof(*t));
  t->speed_preset = speed;
  int pmin, pmax;
  if (speed == 0) { pmin = 4; pmax = 64 < max_block_size ? 64 : max_block_size; }
  else if (speed == 1 && low_quality) { pmin = 4; pmax = 64 < max_block_size ? 64 : max_block_size; }
  else if (speed == 2 && low_quality) { pmin = 4; pmax = 32 < max_block_size ? 32 : max_block_size; }
  else if (speed >= 1 && speed <= 4) { pmin = 4; pmax = 16; }
  else if (speed >= 5 && speed <= 8) { pmin = 8; pmax = 16; }
  else { pmin = 16; pmax = 16; }
  t->partition_min = pmin; t->partition_max = pmax;
  t->complex_prediction_modes = speed <= 1;
  t->sgr_complexity_full = speed <= 2;
  t->encode_bottomup = speed <= 2;
  t->rdo_tx_decision = speed <= 4 && !high_quality;
  t->reduced_tx_set = speed == 4 || speed >= 9;
  t->fine_directional_intra = speed <= 6;
  t->fast_deblock = speed >= 7 && !high_quality;
  t->lrf = low_quality && speed <= 8;
  t->cdef = low_quality && speed <= 9;
  t->inter_tx_split = speed >= 9;
  t->tx_domain_rate = speed >= 10;
  t->tx_domain_distortion = -1;
  t->use_satd_subpel = 0;
  int mts;
  switch (speed) { case 0: mts = 4096; break; case 1: mts = 2048; break; case 2: mts = 1024; break;
                   case 3: mts = 512; break; case 4: mts = 256; break; default: mts = 128; }
  t->min_tile_size = mts * (high_quality ? 2 : 1);
}

/* av1encoder.rs:665-668: tiles = min(threads, w*h / min_tile_size^2) */
size_t orc_tile_target(size_t threads, size_t w, size_t h, size_t min_tile_size) {
  size_t cap = (w * h) / (min_tile_size * min_tile_size);
  return threads < cap ? threads : cap;
}

/* ---- dirtyalpha.rs ---------------------------------------------------------------------- */

/* dirtyalpha.rs:113-124 */
void orc_premultiplied_minmax(uint8_t px, uint8_t alpha, uint8_t out[2]) {
  const uint16_t a = alpha;
  const uint16_t rounded = (uint16_t)((uint16_t)px * a / 255 * 255);
  const uint8_t low = (uint8_t)((rounded + 16) / a);   /* `as u8` on u16: truncation */
  const uint8_t hi = (uint8_t)((rounded + 239) / a);
  out[0] = low < px ? low : px;
  out[1] = hi > px ? hi : px;
}
static inline uint8_t clamp_mm(uint8_t v, const uint8_t mm[2]) { uint8_t t = v > mm[0] ? v : mm[0]; return t < mm[1] ? t : mm[1]; }

/* loop9 0.1.5 semantics [EXT crate]: 3x3 window, rows/columns clamped to the image edge. */
static inline const uint8_t *px_at(const uint8_t *img, size_t w, size_t h, size_t stride, long x, long y) {
  if (x < 0) x = 0; if (y < 0) y = 0;
  if (x >= (long)w) x = (long)w - 1; if (y >= (long)h) y = (long)h - 1;
  return img + ((size_t)y * stride + (size_t)x) * 4;
}

/* dirtyalpha.rs:5-14 */
static inline void weighed_pixel(const uint8_t *p, uint32_t *wgt, uint32_t rgb[3]) {
  if (p[3] == 0) { *wgt = 0; rgb[0] = rgb[1] = rgb[2] = 0; return; }
  const uint32_t w = 256u - p[3];
  *wgt = w; rgb[0] = p[0] * w; rgb[1] = p[1] * w; rgb[2] = p[2] * w;
}

/* dirtyalpha.rs:17-41.  Returns 1 and fills out (w*h*4, tightly packed) when the image has
 * semi-transparent pixels adjacent to fully transparent ones, else 0 (None). */
int orc_blurred_dirty_alpha(const uint8_t *img, size_t w, size_t h, size_t stride, uint8_t *out) {
  uint64_t sum[3] = {0, 0, 0}, weights = 0;
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = px_at(img, w, h, stride, (long)x, (long)y);
      if (c[3] == 255 || c[3] == 0) continue;
      int any0 = 0;
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++)
          if (px_at(img, w, h, stride, (long)x + dx, (long)y + dy)[3] == 0) any0 = 1;
      if (any0) {
        uint32_t wg, rgb[3];
        weighed_pixel(c, &wg, rgb);
        weights += wg; sum[0] += rgb[0]; sum[1] += rgb[1]; sum[2] += rgb[2];
      }
    }
  if (weights == 0) return 0;
  const uint8_t bg[4] = {(uint8_t)(sum[0] / weights), (uint8_t)(sum[1] / weights), (uint8_t)(sum[2] / weights), 0};

  /* bleed_opaque_color, dirtyalpha.rs:45-76 */
  uint8_t *tmp = (uint8_t *)malloc(w * h * 4);
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = px_at(img, w, h, stride, (long)x, (long)y);
      uint8_t *o = tmp + (y * w + x) * 4;
      if (c[3] == 255) { memcpy(o, c, 4); continue; }
      uint32_t ws = 0, s[3] = {0, 0, 0};
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          uint32_t wg, rgb[3];
          weighed_pixel(px_at(img, w, h, stride, (long)x + dx, (long)y + dy), &wg, rgb);
          ws += wg; s[0] += rgb[0]; s[1] += rgb[1]; s[2] += rgb[2];
        }
      if (ws == 0) { memcpy(o, bg, 4); continue; }
      uint8_t avg[3] = {(uint8_t)(s[0] / ws), (uint8_t)(s[1] / ws), (uint8_t)(s[2] / ws)};
      if (c[3] == 0) { o[0] = avg[0]; o[1] = avg[1]; o[2] = avg[2]; o[3] = 0; }
      else {
        for (int k = 0; k < 3; k++) { uint8_t mm[2]; orc_premultiplied_minmax(c[k], c[3], mm); o[k] = clamp_mm(avg[k], mm); }
        o[3] = c[3];
      }
    }
  /* blur_transparent_pixels, dirtyalpha.rs:79-100 */
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = px_at(tmp, w, h, w, (long)x, (long)y);
      uint8_t *o = out + (y * w + x) * 4;
      if (c[3] == 255) { memcpy(o, c, 4); continue; }
      uint16_t s[3] = {0, 0, 0};
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          const uint8_t *n = px_at(tmp, w, h, w, (long)x + dx, (long)y + dy);
          s[0] += n[0]; s[1] += n[1]; s[2] += n[2];
        }
      uint8_t avg[3] = {(uint8_t)(s[0] / 9), (uint8_t)(s[1] / 9), (uint8_t)(s[2] / 9)};
      if (c[3] == 0) { o[0] = avg[0]; o[1] = avg[1]; o[2] = avg[2]; o[3] = 0; }
      else {
        for (int k = 0; k < 3; k++) { uint8_t mm[2]; orc_premultiplied_minmax(c[k], c[3], mm); o[k] = clamp_mm(avg[k], mm); }
        o[3] = c[3];
      }
    }
  free(tmp);
  return 1;
}

/* av1encoder.rs:282-297: AlphaColorMode::Premultiplied branch, exactly as coded
 * (c*255/a in u16 then `as u8` truncation; a==0 or a==255 -> RGBA8::default() = 0,0,0,0). */
void orc_premultiplied_branch(const uint8_t *img, size_t w, size_t h, size_t stride, uint8_t *out) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = img + (y * stride + x) * 4;
      uint8_t *o = out + (y * w + x) * 4;
      if (c[3] == 0 || c[3] == 255) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
      for (int k = 0; k < 3; k++) o[k] = (uint8_t)((uint16_t)((uint16_t)c[k] * 255u / (uint16_t)c[3]));
      o[3] = c[3];
    }
}
