// ravif's in-tree configuration arithmetic, host side (product code).
// Follows /root/reference/ravif/src/av1encoder.rs:526-605 and :665-668; checked bit-exactly against the independent C
// restatement in oracle/ravif_glue.c by tests/test_glue.py.  (The alpha clean-up of dirtyalpha.rs runs on the device:
// alpha_clean.cu.)
#include "ravif_settings.h"

#include <math.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
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

// rayon::current_num_threads() of the global pool (what `threads: None` means, av1encoder.rs:413-415, :666):
// RAYON_NUM_THREADS when set to a positive number, else std::thread::available_parallelism(), which on Linux is the
// size of the affinity mask bounded by the cgroup CPU quota.
extern "C" size_t b200_host_threads(void) {
  if (const char *ev = getenv("RAYON_NUM_THREADS")) { const long v = atol(ev); if (v > 0) return (size_t)v; }
  size_t n = 0;
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = (size_t)CPU_COUNT(&set);
  if (n == 0) { const long v = sysconf(_SC_NPROCESSORS_ONLN); n = v > 0 ? (size_t)v : 1; }
  // cgroup v2 quota: "max 100000" or "<quota> <period>"; v1: cpu.cfs_quota_us / cpu.cfs_period_us
  long long quota = -1, period = 0;
  if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
    char q[32] = {0};
    if (fscanf(f, "%31s %lld", q, &period) == 2 && strcmp(q, "max") != 0) quota = atoll(q);
    fclose(f);
  } else {
    if (FILE *fq = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(fq, "%lld", &quota) != 1) quota = -1; fclose(fq); }
    if (FILE *fp = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(fp, "%lld", &period) != 1) period = 0; fclose(fp); }
  }
  if (quota > 0 && period > 0) n = std::min(n, (size_t)std::max<long long>(quota / period, 1));
  return std::max<size_t>(n, 1);
}

extern "C" size_t b200_tile_target(size_t threads, size_t w, size_t h, size_t min_tile_size) {
  const size_t cap = (w * h) / (min_tile_size * min_tile_size);
  // av1encoder.rs:665-668: tiles = threads.min(cap); threads == 0 is ravif's `None` = the global pool size (:413-415).
  // Like the reference, default-config output therefore depends on the host: callers that need
  // reproducible bytes pass with_num_threads(n); n >= cap gives the most tiles (most device parallelism).
  return std::min(threads == 0 ? b200_host_threads() : threads, cap);
}
