// Interface between the hot path (CUDA kernels; restated on the CPU by oracle/) and the
// host-side symbol packer (av1_pack.cpp): "per-block {mode, tx_type, quantized coeffs,
// reconstructed plane} streamed back for symbol packing" (BASELINE.json north_star).
// Plain C so that the oracle, the CUDA translation units and the C++ host all share it.
#ifndef B200_RECORDS_H_
#define B200_RECORDS_H_
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__
#else
#define B200_HD
#endif

#ifdef __cplusplus
extern "C" {
#endif

// AV1 enumerations (AV1 bitstream numbering)
enum {
  BLOCK_4X4 = 0, BLOCK_4X8, BLOCK_8X4, BLOCK_8X8, BLOCK_8X16, BLOCK_16X8, BLOCK_16X16, BLOCK_16X32,
  BLOCK_32X16, BLOCK_32X32, BLOCK_32X64, BLOCK_64X32, BLOCK_64X64, BLOCK_64X128, BLOCK_128X64, BLOCK_128X128,
  BLOCK_INVALID = 255
};
enum { TX_4X4 = 0, TX_8X8, TX_16X16, TX_32X32, TX_64X64 };
enum {
  DC_PRED = 0, V_PRED, H_PRED, D45_PRED, D135_PRED, D113_PRED, D157_PRED, D203_PRED, D67_PRED,
  SMOOTH_PRED, SMOOTH_V_PRED, SMOOTH_H_PRED, PAETH_PRED, UV_CFL_PRED = 13
};
enum { TX_DCT_DCT = 0, TX_ADST_DCT = 1, TX_DCT_ADST = 2, TX_ADST_ADST = 3, TX_IDTX = 9, TX_V_DCT = 10, TX_H_DCT = 11 };
// searched luma transform types in candidate order (bit k of B200Frame.tx_type_mask enables entry k).  Entries 5 and 6 (the
// 1-D DCTs) exist only in the 7-type intra set: transforms up to 8x8 when reduced_tx_set is off (AV1 spec get_tx_set).
#define B200_TX_TYPES 7
#define B200_TX_TYPE_LIST TX_DCT_DCT, TX_ADST_DCT, TX_DCT_ADST, TX_ADST_ADST, TX_IDTX, TX_V_DCT, TX_H_DCT
// transform class of a type (AV1 spec get_tx_class): 0 = 2-D, 1 = TX_CLASS_HORIZ (H_DCT: identity vertically), 2 = TX_CLASS_VERT
static inline B200_HD int b200_tx_class(int tx_type) { return tx_type == TX_H_DCT ? 1 : (tx_type == TX_V_DCT ? 2 : 0); }
// luma transform types a block of size n may use: the frame's mask, minus the 1-D DCTs where the 5-type set applies
static inline B200_HD int b200_tx_mask_for(int tx_type_mask, int reduced_tx_set, int n) {
  if (n >= 32) return 1;
  return (reduced_tx_set || n >= 16) ? (tx_type_mask & 0x1f) : tx_type_mask;
}

#define B200_MAX_TILE_COLS 64
#define B200_MAX_TILE_ROWS 64
#define B200_SB_COEFS 4096           // 64x64 coefficient slots per plane per superblock

// One entry per 4x4 "mode info" unit (replicated over the whole block).
typedef struct B200Mi {
  uint8_t bsize;     // AV1 BLOCK_SIZE of the block covering this unit
  uint8_t y_mode;    // 0..12
  uint8_t uv_mode;   // 0..13 (13 = CFL)
  uint8_t angle;     // (angle_delta_y + 3) | (angle_delta_uv + 3) << 4
  uint8_t tx_size;   // TX_SIZE of the luma transform (== largest for the block)
  uint8_t tx_type;   // luma transform type (chroma is derived from uv_mode)
  uint8_t flags;     // bit0: skip; bits1..3: cfl joint sign (0..7)
  uint8_t cfl;       // cfl_alpha_u index (low nibble), cfl_alpha_v index (high nibble)
} B200Mi;

// Frame-level description: inputs (geometry, quantiser, tool switches) + decisions made by the
// in-loop-filter search.  Filled by b200av1_frame_setup().
typedef struct B200Frame {
  int32_t width, height;            // visible size
  int32_t bit_depth;                // 8 | 10
  int32_t planes;                   // 3 (4:4:4) or 1 (monochrome)
  int32_t mi_cols, mi_rows;         // 4x4 units, multiple of 2 (8-pixel aligned)
  int32_t sb_cols, sb_rows;         // 64x64 superblocks
  int32_t tile_cols_log2, tile_rows_log2, tile_cols, tile_rows;
  int32_t tile_col_start_sb[B200_MAX_TILE_COLS + 1];
  int32_t tile_row_start_sb[B200_MAX_TILE_ROWS + 1];
  int32_t base_q_idx;               // 1..255
  int32_t dc_q, ac_q;               // quantiser step sizes (table lookups at base_q_idx)
  int32_t reduced_tx_set;
  int32_t enable_cdef, enable_restoration;
  int32_t color_description_present, color_primaries, transfer_characteristics, matrix_coefficients, full_range;
  // encoder search controls (ravif SpeedTweaks, av1encoder.rs:555-605)
  int32_t partition_min, partition_max;     // 4|8|16|32|64 pixels
  int32_t rdo_tx_decision;                  // search transform type per block
  int32_t fine_directional_intra;           // angle_delta search
  int32_t enable_cfl;
  int32_t enable_intra_edge_filter;          // sequence-header tool: edge smoothing / upsampling of directional predictors (AV1 spec 7.11.2.4)
  int32_t lambda_q8;                        // RD multiplier: J = (SSE << 8) + ((lambda_q8 * R_q8) >> 8), R in 1/256 bit
  int32_t lambda_satd_q8;                   // pre-ranking: J = satd * 64 + ((lambda_satd_q8 * R_q8) >> 8)
  int32_t y_mode_mask, uv_mode_mask;        // allowed intra modes (bit = mode index); conformance tests force single modes
  int32_t tx_type_mask;                     // allowed luma transform types: bit0 DCT_DCT 1 ADST_DCT 2 DCT_ADST 3 ADST_ADST 4 IDTX 5 V_DCT 6 H_DCT
  // early termination of the partition search: the children of a block are not analysed when its own RD cost is at
  // most split_gate_bits * lambda ([0]: 8x8 -> 4x4, [1]: 16x16 -> 8x8, [2]: 32x32 -> 16x16; 0 = exhaustive search)
  int32_t split_gate_bits[3];
  // in-loop filter searches (ravif SpeedTweaks): lf_search = !fast_deblock (av1encoder.rs:586): the deblocking level of every
  // plane is searched on the device (B200_LF_CANDIDATES around the closed-form level, squared error against the source on 1/16
  // of the superblocks); lr_full = sgr_complexity_full (:573): all 16 self-guided parameter sets instead of 4
  int32_t lf_search, lr_full;
  // Tune::Psychovisual (ravif always selects it, av1encoder.rs:694): luma blocks of 8x8 and larger are scored per 8x8 with the
  // variance-aware error b200_psy_dist8x8 instead of the plain squared error
  int32_t psy;
  // luma modes that get full trials, in SATD ranking order: 2, or 4 with ravif's complex_prediction_modes (speed <= 1,
  // av1encoder.rs:572,628: rav1e's ComplexAll evaluates more intra candidates per block than Simple)
  int32_t luma_rdo_modes;
  // decisions of the in-loop filter stages
  int32_t lf_level[4];                      // y vertical, y horizontal, u, v
  int32_t lf_sharpness;
  int32_t cdef_damping, cdef_bits;
  int32_t cdef_y_strength[8], cdef_uv_strength[8];   // pri*4 + sec
  int32_t lr_type[3];                       // per plane: 0 RESTORE_NONE, 1 RESTORE_SGRPROJ (64x64 units, lr_unit_shift 0)
} B200Frame;

// Tune::Psychovisual distortion of one 8x8 luma block from its five sums (source s, reconstruction d over the 64 samples).
// rav1e's cdef_dist_kernel [EXT: crates.io rav1e 0.8.1, rdo.rs] scales the squared error by an SSIM-like factor
//   (4033 / 16384) * (svar + dvar + 16384) / sqrt(16265089 + svar * dvar)        (variances at 8-bit scale, x64);
// this is an integer restatement of that published form (Q12 factor, floor square root, factor capped at 4), not rav1e's code.
// floor(sqrt(v)): a floating-point estimate corrected to the exact integer (single precision on the device: the estimate is
// within a few units for v < 2^48 and the loops finish the job; the host uses the double-precision root - same result)
static inline B200_HD uint64_t b200_isqrt64(uint64_t v) {
#ifdef __CUDA_ARCH__
  uint64_t r = (uint64_t)__fsqrt_rn((float)v);
#else
  uint64_t r = (uint64_t)sqrt((double)v);
#endif
  while (r * r > v) r--;
  while ((r + 1) * (r + 1) <= v) r++;
  return r;
}
// floor(num / den) for den != 0, quotients that matter below 2^22 (larger ones only need to stay large: the caller caps them)
static inline B200_HD uint64_t b200_div_small_quot(uint64_t num, uint64_t den) {
#ifdef __CUDA_ARCH__
  const float qf = __fdividef((float)num, (float)den);
  if (qf > 4.0e6f) return 4000000ull;
  uint64_t q = (uint64_t)qf;
  while (q * den > num) q--;
  while ((q + 1) * den <= num) q++;
  return q;
#else
  return num / den;
#endif
}
static inline B200_HD uint32_t b200_psy_dist8x8(uint32_t sum_s, uint32_t sum_d, uint32_t sum_s2, uint32_t sum_d2, uint32_t sum_sd, int bit_depth) {
  const uint64_t sse = (uint64_t)sum_s2 + sum_d2 - 2 * (uint64_t)sum_sd;
  const int cs2 = 2 * (bit_depth - 8);
  const uint64_t svar = ((uint64_t)sum_s2 - (((uint64_t)sum_s * sum_s + 32) >> 6)) >> cs2;
  const uint64_t dvar = ((uint64_t)sum_d2 - (((uint64_t)sum_d * sum_d + 32) >> 6)) >> cs2;
  uint64_t boost = b200_div_small_quot(4033ull * (svar + dvar + 16384), 4 * b200_isqrt64(16265089ull + svar * dvar));      /* Q12 */
  if (boost > 4 * 4096) boost = 4 * 4096;
  const uint64_t d = (sse * boost + 2048) >> 12;
  return d > 0x3fffffffu ? 0x3fffffffu : (uint32_t)d;
}

// Loop restoration decision of one 64x64 restoration unit of one plane (self-guided filter, AV1 spec 5.11.58):
// unit grid = b200_lr_units(height) rows x b200_lr_units(width) columns, row-major.
typedef struct B200LrUnit {
  int8_t use;          // use_sgrproj
  int8_t set;          // lr_sgr_set 0..15
  int8_t xqd0, xqd1;   // LrSgrXqd: -96..31, -32..95
} B200LrUnit;
static inline B200_HD int b200_lr_units(int size) { const int n = (size + 32) >> 6; return n < 1 ? 1 : n; }
// encoder search: candidate parameter sets and the static cost (1/256 bit) of the use_sgrproj flag (default CDF 16855)
#define B200_LR_CANDIDATES 4
#define B200_LR_SETS 2, 6, 11, 14
#define B200_LR_CANDIDATES_FULL 16       /* lr_full: every set of av1_sgr_params */
// deblocking level search: candidate k = clamp((L0 * mult[k] + 2) >> 2, 0, 63), L0 = closed-form level; ties keep the lowest k;
// evaluated on every fourth superblock row and column (1/16 of the frame: the level is one number per plane and frame)
#define B200_LF_CANDIDATES 4
#define B200_LF_MULTS 4, 2, 6, 0
static inline B200_HD int b200_lf_candidate(int l0, int k) {
  const int mult[B200_LF_CANDIDATES] = {B200_LF_MULTS};
  const int v = (l0 * mult[k] + 2) >> 2;
  return v > 63 ? 63 : v;
}
static inline B200_HD int b200_lf_sampled(int sbr, int sbc) { return !((sbr | sbc) & 3); }
#define B200_LR_FLAG_COST0 246
#define B200_LR_FLAG_COST1 267

// Decision of the open-loop analysis for one block (analysis stage -> partition decision -> reconstruction).
typedef struct B200BlockDec {
  int64_t cost;        // J = SSE*256 + lambda*R of the best (mode, transform type) per plane, + skip flag cost
  uint8_t y_mode, tx_type, uv_mode;
  uint8_t all_zero;    // every plane quantised to zero in the analysis (the block would be coded as "skip")
  uint8_t angle_y, angle_uv;   // angle_delta + 3 of the chosen luma / chroma mode (3 = no delta)
  uint8_t cfl_js, cfl_alpha;   // uv_mode == UV_CFL_PRED: joint sign (0..7), alpha indices (u low nibble, v high nibble)
} B200BlockDec;
// SATD table entry per block (luma): 13 mode SATDs + one word packing the refined directional mode and its
// angle delta: (mode | (delta + 3) << 8), mode = 255 when no directional mode was refined.
#define B200_SATD_WORDS 14
// analysis levels: 0: 4x4, 1: 8x8, 2: 16x16, 3: 32x32 (block size 4 << lv)
#define B200_LEVELS 4
// Early termination of the partition search: the four 4x4 children of an 8x8 block whose analysis quantised to
// all-zero are not analysed; their cost is this sentinel so the split can never win.
#define B200_COST_PRUNED ((int64_t)1 << 60)
// Number of analysed blocks per row / column at level lv (0: 4x4, 1: 8x8, 2: 16x16); blocks that stick out of the
// (8-pixel aligned) coded frame are not analysed.
static inline B200_HD int b200_level_cols(const struct B200Frame *f, int lv);
static inline B200_HD int b200_level_rows(const struct B200Frame *f, int lv);

// Static rate model used by the RDO (costs in 1/256 bit, derived from the default CDFs of the
// frame's quantiser class, no adaptation).  Built on the host by b200av1_build_rate_tables().
typedef struct B200RateTables {
  uint16_t partition[20][10];
  uint16_t skip[3][2];
  uint16_t y_mode[5][5][13];
  uint16_t uv_mode[2][13][14];
  uint16_t angle_delta[8][7];
  uint16_t cfl_sign[8];
  uint16_t cfl_alpha[6][16];
  uint16_t intra_tx[2][4][13][16];          // [0: 7-type set, 1: 5-type set][tx size][y mode][tx_type]
  uint16_t txb_skip[5][13][2];
  uint16_t eob_pt[4][2][2][12];             // [TX_4X4..TX_32X32][plane type][0: 2-D transform class, 1: 1-D class][eob_pt - 1]
  uint16_t eob_extra[5][2][9][2];
  uint16_t dc_sign[2][3][2];
  uint16_t coeff_base_eob[5][2][4][3];
  uint16_t coeff_base[5][2][42][4];
  uint16_t coeff_br[4][2][21][13];          // cost of base_range 0..12 (12 = 12 or more, golomb extra not included)
} B200RateTables;

static inline B200_HD int b200_level_cols(const B200Frame *f, int lv) { return f->mi_cols >> lv; }
static inline B200_HD int b200_level_rows(const B200Frame *f, int lv) { return f->mi_rows >> lv; }

// Top-down pruning of the analysis (same rule in the oracle and in every device kernel): block (bx, by) of level
// lv (0: 4x4, 1: 8x8, 2: 16x16) is not analysed when its parent at level lv + 1 was itself pruned, quantised to all-zero in
// every plane, or is already cheap enough (split_gate_bits).  `parent` = the decision array of level lv + 1.
static inline B200_HD int b200_pruned_by_parent(const B200Frame *f, int lv, int bx, int by, const B200BlockDec *parent) {
  if (lv >= B200_LEVELS - 1) return 0;
  const int pn = 8 << lv;
  const int pmin8 = f->partition_min < 8 ? f->partition_min : 8;
  if (pn < pmin8 || pn > f->partition_max) return 0;                  /* the parent level is not analysed */
  const int pc = b200_level_cols(f, lv + 1), pr = b200_level_rows(f, lv + 1);
  if ((bx >> 1) >= pc || (by >> 1) >= pr) return 0;                   /* parent sticks out of the frame: forced split */
  const B200BlockDec *d = parent + (size_t)(by >> 1) * pc + (bx >> 1);
  if (d->cost >= B200_COST_PRUNED || d->all_zero) return 1;
  const int gate = f->split_gate_bits[lv];
  return gate > 0 && d->cost <= (int64_t)gate * f->lambda_q8;
}

// z-order index (0..255) of the 4x4 unit (r, c) inside its 64x64 superblock, r,c in 0..15.
static inline B200_HD int b200_spread4(int v) {   // abcd -> 0a0b0c0d
  v = (v | (v << 2)) & 0x33;
  return (v | (v << 1)) & 0x55;
}
static inline B200_HD int b200_zidx(int r, int c) { return b200_spread4(c) | (b200_spread4(r) << 1); }
// Is the 4x4 unit (ur, uc) reconstructed before the block whose top-left unit is (br, bc)?  Coding order is
// static: superblocks in raster order inside the tile [tr0,tr1) x [tc0,tc1), z-order inside a superblock; any block
// containing the unit lies entirely before or after the current block.  (AV1 top-right / bottom-left availability.)
static inline B200_HD int b200_unit_before(int ur, int uc, int br, int bc, int tr0, int tr1, int tc0, int tc1) {
  if (ur < tr0 || ur >= tr1 || uc < tc0 || uc >= tc1) return 0;
  const int usr = ur >> 4, usc = uc >> 4, bsr = br >> 4, bsc = bc >> 4;
  if (usr != bsr) return usr < bsr;
  if (usc != bsc) return usc < bsc;
  return b200_zidx(ur & 15, uc & 15) < b200_zidx(br & 15, bc & 15);
}
// First coefficient slot of the transform block whose top-left 4x4 unit is (mi_row, mi_col).
// Coefficients of a block are stored in SCAN ORDER, contiguous, in its z-order slot range.
static inline B200_HD int64_t b200_coef_offset(const B200Frame *f, int mi_row, int mi_col) {
  const int sb = (mi_row >> 4) * f->sb_cols + (mi_col >> 4);
  return (int64_t)sb * B200_SB_COEFS + b200_zidx(mi_row & 15, mi_col & 15) * 16;
}

#ifdef __cplusplus
}
#endif
#endif  // B200_RECORDS_H_
