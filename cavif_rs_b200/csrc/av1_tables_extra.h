// AV1 normative constants that are anonymous in the libaom binary used by
// tools/extract_aom_tables.py, restated by hand.  tests/test_tables.py cross-checks each of
// them by byte-pattern search in libaom's / dav1d's .rodata and they are exercised through
// dav1d decode-match tests.  CDF convention as in av1_tables.h: 32768 - cumulative, plus one
// trailing adaptation-counter slot.
#ifndef AV1_TABLES_EXTRA_H_
#define AV1_TABLES_EXTRA_H_
#include <stdint.h>
#ifndef AV1_TABLE
#define AV1_TABLE(type, name, dims) static const type name dims
#endif

#define ICDF(x) (uint16_t)(32768 - (x))

AV1_TABLE(uint16_t, av1_default_skip_cdf, [3][3]) = {
  {ICDF(31671), 0, 0}, {ICDF(16515), 0, 0}, {ICDF(4576), 0, 0},
};

AV1_TABLE(uint16_t, av1_default_angle_delta_cdf, [8][8]) = {
  {ICDF(2180), ICDF(5032), ICDF(7567), ICDF(22776), ICDF(26989), ICDF(30217), 0, 0},
  {ICDF(2301), ICDF(5608), ICDF(8801), ICDF(23487), ICDF(26974), ICDF(30330), 0, 0},
  {ICDF(3780), ICDF(11018), ICDF(13699), ICDF(19354), ICDF(23083), ICDF(31286), 0, 0},
  {ICDF(4581), ICDF(11226), ICDF(15147), ICDF(17138), ICDF(21834), ICDF(28397), 0, 0},
  {ICDF(1737), ICDF(10927), ICDF(14509), ICDF(19588), ICDF(22745), ICDF(28823), 0, 0},
  {ICDF(2664), ICDF(10176), ICDF(12485), ICDF(17650), ICDF(21600), ICDF(30495), 0, 0},
  {ICDF(2240), ICDF(11096), ICDF(15453), ICDF(20341), ICDF(22561), ICDF(28917), 0, 0},
  {ICDF(3605), ICDF(10428), ICDF(12459), ICDF(17676), ICDF(21244), ICDF(30655), 0, 0},
};

AV1_TABLE(uint16_t, av1_default_cfl_sign_cdf, [9]) = {
  ICDF(1418), ICDF(2123), ICDF(13340), ICDF(18405), ICDF(26972), ICDF(28343), ICDF(32294), 0, 0,
};

AV1_TABLE(uint16_t, av1_default_cfl_alpha_cdf, [6][17]) = {
  {ICDF(7637), ICDF(20719), ICDF(31401), ICDF(32481), ICDF(32657), ICDF(32688), ICDF(32692), ICDF(32696),
   ICDF(32700), ICDF(32704), ICDF(32708), ICDF(32712), ICDF(32716), ICDF(32720), ICDF(32724), 0, 0},
  {ICDF(14365), ICDF(23603), ICDF(28135), ICDF(31168), ICDF(32167), ICDF(32395), ICDF(32487), ICDF(32573),
   ICDF(32620), ICDF(32647), ICDF(32668), ICDF(32672), ICDF(32676), ICDF(32680), ICDF(32684), 0, 0},
  {ICDF(11532), ICDF(22380), ICDF(28445), ICDF(31360), ICDF(32349), ICDF(32523), ICDF(32584), ICDF(32649),
   ICDF(32673), ICDF(32677), ICDF(32681), ICDF(32685), ICDF(32689), ICDF(32693), ICDF(32697), 0, 0},
  {ICDF(26990), ICDF(31402), ICDF(32282), ICDF(32571), ICDF(32692), ICDF(32696), ICDF(32700), ICDF(32704),
   ICDF(32708), ICDF(32712), ICDF(32716), ICDF(32720), ICDF(32724), ICDF(32728), ICDF(32732), 0, 0},
  {ICDF(17248), ICDF(26058), ICDF(28904), ICDF(30608), ICDF(31305), ICDF(31877), ICDF(32126), ICDF(32321),
   ICDF(32394), ICDF(32464), ICDF(32516), ICDF(32560), ICDF(32576), ICDF(32593), ICDF(32622), 0, 0},
  {ICDF(14738), ICDF(21678), ICDF(25779), ICDF(27901), ICDF(29024), ICDF(30302), ICDF(30980), ICDF(31843),
   ICDF(32144), ICDF(32413), ICDF(32520), ICDF(32594), ICDF(32622), ICDF(32656), ICDF(32660), 0, 0},
};

// loop-restoration symbols
AV1_TABLE(uint16_t, av1_default_switchable_restore_cdf, [4]) = {ICDF(9413), ICDF(22581), 0, 0};
AV1_TABLE(uint16_t, av1_default_wiener_restore_cdf, [3]) = {ICDF(11570), 0, 0};
AV1_TABLE(uint16_t, av1_default_sgrproj_restore_cdf, [3]) = {ICDF(16855), 0, 0};

// Context class of a neighbouring intra mode for the key-frame y-mode CDF.
AV1_TABLE(uint8_t, av1_intra_mode_context, [13]) = {0, 1, 2, 3, 4, 4, 4, 4, 3, 0, 1, 2, 0};

// Smooth-predictor weights for block dimension 4, 8, 16, 32, 64 (offset = dimension - 4 ... see sm_w()).
AV1_TABLE(uint8_t, av1_sm_weights_4, [4]) = {255, 149, 85, 64};
AV1_TABLE(uint8_t, av1_sm_weights_8, [8]) = {255, 197, 146, 105, 73, 50, 37, 32};
AV1_TABLE(uint8_t, av1_sm_weights_16, [16]) = {255, 225, 196, 170, 145, 123, 102, 84, 68, 54, 43, 33, 26, 20, 17, 16};
AV1_TABLE(uint8_t, av1_sm_weights_32, [32]) = {255, 240, 225, 210, 196, 182, 169, 157, 145, 133, 122, 111, 101, 92, 83, 74,
                                               66,  59,  52,  45,  39,  34,  29,  25,  21,  17,  14,  12,  10,  9,  8,  8};
AV1_TABLE(uint8_t, av1_sm_weights_64, [64]) = {
    255, 248, 240, 233, 225, 218, 210, 203, 196, 189, 182, 176, 169, 163, 156, 150, 144, 138, 133, 127, 121, 116,
    111, 106, 101, 96,  91,  86,  82,  77,  73,  69,  65,  61,  57,  54,  50,  47,  44,  41,  38,  35,  32,  29,
    27,  25,  22,  20,  18,  16,  15,  13,  12,  10,  9,   8,   7,   6,   6,   5,   5,   4,   4,   4};

#endif  // AV1_TABLES_EXTRA_H_
