// ravif's in-tree configuration arithmetic and alpha clean-up, host side (product code).
// Follows /root/reference/ravif/src/av1encoder.rs:526-605, :665-668, :282-297 and
// /root/reference/ravif/src/dirtyalpha.rs:5-124; checked bit-exactly against the independent C
// restatement in oracle/ravif_glue.c by tests/test_glue.py.
#include "ravif_settings.h"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

extern "C" uint8_t b200_quality_to_quantizer(float quality) {
  const float q = quality / 100.f;
  const float x = q >= 0.82f ? (1.f - q) * 2.6f : (q > 0.25f ? fmaf(q, -0.5f, 1.f - 0.125f) : 1.f - q);
  const float v = roundf(x * 255.f);
  return (uint8_t)(v <= 0.f ? 0.f : (v >= 255.f ? 255.f : v));
}

extern "C" void b200_speed_tweaks(int speed, int quantizer, B200SpeedTweaks *t) {
  const bool low_quality = quantizer < b200_quality_to_quantizer(55.f);
  const bool high_quality = quantizer > b200_quality_to_quantizer(80.f);
  const int max_block_size = high_quality ? 16 : 64;
  memset(t, 0, sizeof(*t));
  t->speed_preset = speed;
  if (speed == 0 || (speed == 1 && low_quality)) { t->partition_min = 4; t->partition_max = std::min(64, max_block_size); }
  else if (speed == 2 && low_quality) { t->partition_min = 4; t->partition_max = std::min(32, max_block_size); }
  else if (speed >= 1 && speed <= 4) { t->partition_min = 4; t->partition_max = 16; }
  else if (speed >= 5 && speed <= 8) { t->partition_min = 8; t->partition_max = 16; }
  else { t->partition_min = 16; t->partition_max = 16; }
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
  static const int mts[5] = {4096, 2048, 1024, 512, 256};
  t->min_tile_size = (speed <= 4 ? mts[speed] : 128) * (high_quality ? 2 : 1);
}

extern "C" size_t b200_tile_target(size_t threads, size_t w, size_t h, size_t min_tile_size) {
  const size_t cap = (w * h) / (min_tile_size * min_tile_size);
  // threads == 0 is ravif's `None` ("all threads of the pool"); the device always has more lanes than the
  // cap, so None resolves to the cap itself (state `threads` explicitly for byte parity).
  return threads == 0 ? cap : std::min(threads, cap);
}

namespace {
inline void minmax(uint8_t px, uint8_t alpha, uint8_t mm[2]) {   // dirtyalpha.rs:113-124
  const uint16_t a = alpha;
  const uint16_t rounded = (uint16_t)((uint16_t)px * a / 255 * 255);
  const uint8_t low = (uint8_t)((rounded + 16) / a), hi = (uint8_t)((rounded + 239) / a);
  mm[0] = std::min(low, px);
  mm[1] = std::max(hi, px);
}
inline uint8_t clampmm(uint8_t v, const uint8_t mm[2]) { return std::min(std::max(v, mm[0]), mm[1]); }
struct View {
  const uint8_t *p; size_t w, h, stride;
  const uint8_t *at(long x, long y) const {
    x = std::min<long>(std::max<long>(x, 0), (long)w - 1);
    y = std::min<long>(std::max<long>(y, 0), (long)h - 1);
    return p + ((size_t)y * stride + (size_t)x) * 4;
  }
};
inline uint32_t weighed(const uint8_t *p, uint32_t rgb[3]) {   // dirtyalpha.rs:5-14
  if (p[3] == 0) { rgb[0] = rgb[1] = rgb[2] = 0; return 0; }
  const uint32_t w = 256u - p[3];
  rgb[0] = p[0] * w; rgb[1] = p[1] * w; rgb[2] = p[2] * w;
  return w;
}
}  // namespace

extern "C" int b200_blurred_dirty_alpha(const uint8_t *rgba, size_t w, size_t h, size_t stride_px, uint8_t *out) {
  const View img{rgba, w, h, stride_px};
  uint64_t sum[3] = {0, 0, 0}, weights = 0;
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = img.at((long)x, (long)y);
      if (c[3] == 255 || c[3] == 0) continue;
      bool near0 = false;
      for (int dy = -1; dy <= 1 && !near0; dy++)
        for (int dx = -1; dx <= 1; dx++)
          if (img.at((long)x + dx, (long)y + dy)[3] == 0) { near0 = true; break; }
      if (!near0) continue;
      uint32_t rgb[3];
      weights += weighed(c, rgb);
      sum[0] += rgb[0]; sum[1] += rgb[1]; sum[2] += rgb[2];
    }
  if (weights == 0) return 0;
  const uint8_t bg[4] = {(uint8_t)(sum[0] / weights), (uint8_t)(sum[1] / weights), (uint8_t)(sum[2] / weights), 0};
  std::vector<uint8_t> tmp(w * h * 4);
  for (size_t y = 0; y < h; y++)      // bleed_opaque_color, dirtyalpha.rs:45-76
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = img.at((long)x, (long)y);
      uint8_t *o = &tmp[(y * w + x) * 4];
      if (c[3] == 255) { memcpy(o, c, 4); continue; }
      uint32_t ws = 0, s[3] = {0, 0, 0};
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          uint32_t rgb[3];
          ws += weighed(img.at((long)x + dx, (long)y + dy), rgb);
          s[0] += rgb[0]; s[1] += rgb[1]; s[2] += rgb[2];
        }
      if (ws == 0) { memcpy(o, bg, 4); continue; }
      for (int k = 0; k < 3; k++) {
        const uint8_t avg = (uint8_t)(s[k] / ws);
        if (c[3] == 0) o[k] = avg;
        else { uint8_t mm[2]; minmax(c[k], c[3], mm); o[k] = clampmm(avg, mm); }
      }
      o[3] = c[3];
    }
  const View t2{tmp.data(), w, h, w};
  for (size_t y = 0; y < h; y++)      // blur_transparent_pixels, dirtyalpha.rs:79-100
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = t2.at((long)x, (long)y);
      uint8_t *o = out + (y * w + x) * 4;
      if (c[3] == 255) { memcpy(o, c, 4); continue; }
      uint16_t s[3] = {0, 0, 0};
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          const uint8_t *n = t2.at((long)x + dx, (long)y + dy);
          s[0] += n[0]; s[1] += n[1]; s[2] += n[2];
        }
      for (int k = 0; k < 3; k++) {
        const uint8_t avg = (uint8_t)(s[k] / 9);
        if (c[3] == 0) o[k] = avg;
        else { uint8_t mm[2]; minmax(c[k], c[3], mm); o[k] = clampmm(avg, mm); }
      }
      o[3] = c[3];
    }
  return 1;
}

extern "C" void b200_premultiplied_branch(const uint8_t *rgba, size_t w, size_t h, size_t stride_px, uint8_t *out) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const uint8_t *c = rgba + (y * stride_px + x) * 4;
      uint8_t *o = out + (y * w + x) * 4;
      if (c[3] == 0 || c[3] == 255) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
      for (int k = 0; k < 3; k++) o[k] = (uint8_t)((uint16_t)((uint16_t)c[k] * 255u / (uint16_t)c[3]));
      o[3] = c[3];
    }
}
