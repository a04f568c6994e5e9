// ravif's in-tree configuration arithmetic, host side (product code).
// Follows /root/reference/ravif/src/av1encoder.rs:526-605 and :665-668; checked bit-exactly against the independent C
// restatement in oracle/ravif_glue.c by tests/test_glue.py.  (The alpha clean-up of dirtyalpha.rs runs on the device:
// alpha_clean.cu.)
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
