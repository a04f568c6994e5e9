// Frame set-up shared by the device path and the packer: geometry, uniform tile layout,
// quantiser steps, RD multipliers and the static rate tables of the RDO.
//
// Tile count follows ravif: tiles = min(threads, w*h / min_tile_size^2)
// (/root/reference/ravif/src/av1encoder.rs:665-668); the split into a uniform
// tile_cols_log2 x tile_rows_log2 grid happens inside rav1e [EXT] and is restated here as
// "split the longer tile side until the target count is reached".
#include <math.h>
#include <string.h>
#include <algorithm>

#include "av1_pack.h"
#include "av1_tables.h"
#include "av1_tables_extra.h"

namespace {
int tile_log2(int blk, int target) {
  int k = 0;
  while ((blk << k) < target) k++;
  return k;
}
uint32_t isqrt64(uint64_t v) {
  uint64_t r = (uint64_t)sqrt((double)v);
  while (r * r > v) r--;
  while ((r + 1) * (r + 1) <= v) r++;
  return (uint32_t)r;
}
void layout(int sbs, int lg, int *count, int32_t *starts) {
  const int w = (sbs + (1 << lg) - 1) >> lg;
  int i = 0;
  for (int s = 0; s < sbs; s += w) starts[i++] = s;
  starts[i] = sbs;
  *count = i;
}
}  // namespace

extern "C" int b200av1_frame_setup(B200Frame *f, int width, int height, int bit_depth, int planes, int quantizer,
                                   int tiles_target) {
  if (!f || width < 1 || height < 1 || width > 65536 || height > 65536) return 4;
  if ((bit_depth != 8 && bit_depth != 10) || (planes != 1 && planes != 3)) return 4;
  memset(f, 0, sizeof(*f));
  f->width = width; f->height = height; f->bit_depth = bit_depth; f->planes = planes;
  f->mi_cols = 2 * ((width + 7) >> 3);
  f->mi_rows = 2 * ((height + 7) >> 3);
  f->sb_cols = (f->mi_cols + 15) >> 4;
  f->sb_rows = (f->mi_rows + 15) >> 4;
  // --- tiles (uniform spacing)
  const int max_w_sb = 4096 >> 6, max_area_sb = (4096 * 2304) >> 12;
  const int min_log2_cols = tile_log2(max_w_sb, f->sb_cols);
  const int max_log2_cols = tile_log2(1, std::min(f->sb_cols, 64));
  const int max_log2_rows = tile_log2(1, std::min(f->sb_rows, 64));
  const int min_log2_tiles = std::max(min_log2_cols, tile_log2(max_area_sb, f->sb_rows * f->sb_cols));
  int cl = min_log2_cols, rl = std::max(min_log2_tiles - cl, 0);
  rl = std::min(rl, max_log2_rows);
  for (;;) {
    layout(f->sb_cols, cl, &f->tile_cols, f->tile_col_start_sb);
    layout(f->sb_rows, rl, &f->tile_rows, f->tile_row_start_sb);
    if (f->tile_cols * f->tile_rows >= tiles_target) break;
    const int tw = (f->sb_cols + (1 << cl) - 1) >> cl, th = (f->sb_rows + (1 << rl) - 1) >> rl;
    const bool can_c = cl < max_log2_cols && tw > 1, can_r = rl < max_log2_rows && th > 1;
    if (!can_c && !can_r) break;
    if ((tw >= th && can_c) || !can_r) cl++; else rl++;
  }
  f->tile_cols_log2 = cl; f->tile_rows_log2 = rl;
  // --- quantiser: ravif's 0..255 "quantizer" (av1encoder.rs:526-530) is used directly as base_q_idx
  // (rav1e additionally applies a log-domain key-frame offset [EXT, unpinned]).  qidx 0 is lossless in AV1
  // and is not supported by this encoder, so clamp to 1.
  const int qi = std::min(std::max(quantizer, 1), 255);
  f->base_q_idx = qi;
  f->dc_q = bit_depth == 8 ? av1_dc_qlookup_8[qi] : av1_dc_qlookup_10[qi];
  f->ac_q = bit_depth == 8 ? av1_ac_qlookup_8[qi] : av1_ac_qlookup_10[qi];
  // lambda = ln2/6 * (ac_q/8)^2 per bit of rate, in SSE units  ->  Q8
  const int64_t q2 = (int64_t)f->ac_q * f->ac_q;
  f->lambda_q8 = (int32_t)std::min<int64_t>((q2 * 473 + 512) >> 10, 0x7fffffff);
  f->lambda_satd_q8 = (int32_t)isqrt64((uint64_t)f->lambda_q8 << 8);
  f->reduced_tx_set = 1;
  f->partition_min = 4; f->partition_max = 16;
  f->luma_rdo_modes = 2;
  f->rdo_tx_decision = 1;
  f->fine_directional_intra = 1;
  f->enable_cfl = 1;
  f->y_mode_mask = 0x1fff; f->uv_mode_mask = 0x1fff; f->tx_type_mask = 0x7f;
  // In-loop filter strengths: closed-form picks from the quantiser (encoder freedom; the filters themselves
  // are normative).  Deblock level: libaom's "pick from q" fit; CDEF: one frame-wide strength pair.
  {
    const int q = f->ac_q;
    int lvl = bit_depth == 8 ? (int)(((int64_t)q * 17563 - 421574 + (1 << 17)) >> 18)
                             : (int)(((int64_t)q * 20723 + 4060632 + (1 << 19)) >> 20);
    lvl = std::min(std::max(lvl, 0), 63);
    f->lf_level[0] = f->lf_level[1] = f->lf_level[2] = f->lf_level[3] = lvl;
    f->lf_sharpness = 0;
    const int q8 = q >> (bit_depth - 8);
    // CDEF: four candidate strength pairs (cdef_bits = 2) scaled from the quantiser; index 0 = off.  The
    // per-superblock index is chosen on the device by squared error against the source (cdef_search).
    const int yp = std::min(std::max((q8 * 7 + 512) >> 10, 1), 15);
    const int up = std::min(std::max((q8 * 7 + 1024) >> 11, 1), 15);
    f->cdef_damping = 3 + (q8 >= 128) + (q8 >= 512);
    f->cdef_bits = 2;
    const int ypri[4] = {0, yp, std::min(2 * yp, 15), std::min(3 * yp + 1, 15)}, ysec[4] = {0, 0, 1, 2};
    const int upri[4] = {0, up, std::min(2 * up, 15), std::min(3 * up + 1, 15)}, usec[4] = {0, 0, 1, 1};
    for (int i = 0; i < 4; i++) { f->cdef_y_strength[i] = ypri[i] * 4 + ysec[i]; f->cdef_uv_strength[i] = upri[i] * 4 + usec[i]; }
    f->enable_cdef = 1;
  }
  f->color_description_present = 0; f->color_primaries = 2; f->transfer_characteristics = 2;
  f->matrix_coefficients = 2; f->full_range = 1;
  return 0;
}

namespace {
inline uint16_t cost_of(int p15) {   // p15: probability in 1/32768 units
  if (p15 < 1) p15 = 1;
  double c = -log2((double)p15 / 32768.0) * 256.0;
  int v = (int)(c + 0.5);
  return (uint16_t)std::min(v, 65535);
}
// costs of all symbols of an inverse CDF with n symbols (respecting the coder's minimum probability)
void cdf_costs(const uint16_t *icdf, int n, uint16_t *out) {
  for (int s = 0; s < n; s++) {
    const int hi = s > 0 ? icdf[s - 1] : 32768;
    int p = hi - icdf[s];
    p = std::max(p, 4 * 4);   // EC_MIN_PROB floor (approx.)
    out[s] = cost_of(p);
  }
}
}  // namespace

extern "C" void b200av1_build_rate_tables(int base_q_idx, B200RateTables *t) {
  const int q = base_q_idx <= 20 ? 0 : base_q_idx <= 60 ? 1 : base_q_idx <= 120 ? 2 : 3;
  memset(t, 0, sizeof(*t));
  for (int c = 0; c < 20; c++) cdf_costs(av1_default_partition_cdf[c], c < 4 ? 4 : (c >= 16 ? 8 : 10), t->partition[c]);
  for (int c = 0; c < 3; c++) cdf_costs(av1_default_skip_cdf[c], 2, t->skip[c]);
  for (int a = 0; a < 5; a++) for (int l = 0; l < 5; l++) cdf_costs(av1_default_kf_y_mode_cdf[a][l], 13, t->y_mode[a][l]);
  for (int m = 0; m < 13; m++) { cdf_costs(av1_default_uv_mode_cdf[0][m], 13, t->uv_mode[0][m]); cdf_costs(av1_default_uv_mode_cdf[1][m], 14, t->uv_mode[1][m]); }
  for (int d = 0; d < 8; d++) cdf_costs(av1_default_angle_delta_cdf[d], 7, t->angle_delta[d]);
  cdf_costs(av1_default_cfl_sign_cdf, 8, t->cfl_sign);
  for (int c = 0; c < 6; c++) cdf_costs(av1_default_cfl_alpha_cdf[c], 16, t->cfl_alpha[c]);
  for (int set = 0; set < 2; set++) {
    const int eset = set == 0 ? 1 : 2, set_type = set == 0 ? 3 : 2, nsym = set == 0 ? 7 : 5;
    for (int txs = 0; txs < 4; txs++)
      for (int m = 0; m < 13; m++) {
        uint16_t c[16];
        cdf_costs(av1_default_intra_ext_tx_cdf[eset][txs][m], nsym, c);
        for (int tt = 0; tt < 16; tt++) {
          // valid types of the set: inverse map lists them
          bool in_set = false;
          for (int s = 0; s < nsym; s++) if (av1_ext_tx_inv_tbl[set_type][s] == tt) in_set = true;
          t->intra_tx[set][txs][m][tt] = in_set ? c[av1_ext_tx_ind_tbl[set_type][tt]] : 0xffff;
        }
      }
  }
  for (int s = 0; s < 5; s++) for (int c = 0; c < 13; c++) cdf_costs(av1_default_txb_skip_cdf[q][s][c], 2, t->txb_skip[s][c]);
  for (int p = 0; p < 2; p++) {
    for (int c = 0; c < 2; c++) {
      cdf_costs(av1_default_eob_multi16_cdf[q][p][c], 5, t->eob_pt[0][p][c]);
      cdf_costs(av1_default_eob_multi64_cdf[q][p][c], 7, t->eob_pt[1][p][c]);
      cdf_costs(av1_default_eob_multi256_cdf[q][p][c], 9, t->eob_pt[2][p][c]);
      cdf_costs(av1_default_eob_multi1024_cdf[q][p][c], 11, t->eob_pt[3][p][c]);
    }
  }
  for (int s = 0; s < 5; s++) for (int p = 0; p < 2; p++) for (int c = 0; c < 9; c++)
    cdf_costs(av1_default_eob_extra_cdf[q][s][p][c], 2, t->eob_extra[s][p][c]);
  for (int p = 0; p < 2; p++) for (int c = 0; c < 3; c++) cdf_costs(av1_default_dc_sign_cdf[q][p][c], 2, t->dc_sign[p][c]);
  for (int s = 0; s < 5; s++) for (int p = 0; p < 2; p++) {
    for (int c = 0; c < 4; c++) cdf_costs(av1_default_coeff_base_eob_cdf[q][s][p][c], 3, t->coeff_base_eob[s][p][c]);
    for (int c = 0; c < 42; c++) cdf_costs(av1_default_coeff_base_cdf[q][s][p][c], 4, t->coeff_base[s][p][c]);
  }
  for (int s = 0; s < 4; s++) for (int p = 0; p < 2; p++) for (int c = 0; c < 21; c++) {
    uint16_t k[4];
    cdf_costs(av1_default_coeff_br_cdf[q][s][p][c], 4, k);
    for (int br = 0; br <= 12; br++) {
      int cost = 0, rem = br;
      for (int idx = 0; idx < 12; idx += 3) {
        const int sym = std::min(rem, 3);
        cost += k[sym];
        if (sym < 3) break;
        rem -= 3;
      }
      t->coeff_br[s][p][c][br] = (uint16_t)std::min(cost, 65535);
    }
  }
}
