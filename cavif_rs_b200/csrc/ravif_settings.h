// Host-side restatement (product code) of ravif's in-tree configuration arithmetic.
#ifndef B200_RAVIF_SETTINGS_H_
#define B200_RAVIF_SETTINGS_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct B200SpeedTweaks {
  int speed_preset;
  int fast_deblock, reduced_tx_set, tx_domain_rate, encode_bottomup, rdo_tx_decision;
  int cdef, lrf, sgr_complexity_full, inter_tx_split, fine_directional_intra, complex_prediction_modes;
  int partition_min, partition_max;
  int min_tile_size;
} B200SpeedTweaks;
uint8_t b200_quality_to_quantizer(float quality);                       // av1encoder.rs:526-530
void b200_speed_tweaks(int speed, int quantizer, B200SpeedTweaks *t);    // av1encoder.rs:555-605
size_t b200_host_threads(void);                                        // rayon::current_num_threads() (:413-415)
size_t b200_tile_target(size_t threads /*0 = None*/, size_t w, size_t h, size_t min_tile_size);   // :665-668
#ifdef __cplusplus
}
#endif
#endif
