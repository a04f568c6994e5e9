/* ORACLE (test infrastructure, never linked into the product path).
 *
 * CPU restatement of the device hot path: per-tile AV1 intra rate-distortion search
 * (intra prediction, forward transform, quantisation, reconstruction, distortion,
 * partition / mode / transform-type decisions) producing the block records of
 * cavif_rs_b200/csrc/records.h plus the reconstructed planes.
 *
 * In the reference this is rav1e's encode_tile / rdo_mode_decision / encode_tx_block
 * [EXT: crates.io rav1e 0.8.1, not on disk; driven only from
 * /root/reference/ravif/src/av1encoder.rs:749-771 with the tool switches of :555-605].
 * PARITY UNPINNED against rav1e (no source, no vectors - SURVEY.md section 0.3).  What IS
 * pinned: every normative stage (predictors, dequantisation, inverse transforms,
 * reconstruction) sample-exactly by dav1d 1.5.3 decoding the packed bitstream, and
 * predictors/transforms against libaom 3.13.1's C reference functions (tests/).
 *
 * The CUDA kernels in cavif_rs_b200/csrc/rdo.cu must reproduce this file bit-exactly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../cavif_rs_b200/csrc/records.h"
#include "../cavif_rs_b200/csrc/av1_tables.h"
#include "../cavif_rs_b200/csrc/av1_tables_extra.h"

void orc_fwd_txfm2d(const int16_t *resid, int stride, int32_t *coef, int n, int tx_type);
void orc_inv_txfm2d_add(const int32_t *coef, uint16_t *dst, int stride, int n, int tx_type, int bd);

typedef struct {
  const B200Frame *f;
  const B200RateTables *rt;
  uint16_t *src[3];   /* padded source planes, stride = mi_cols*4 */
  uint16_t *rec[3];   /* reconstruction (before in-loop filters) */
  int stride;
  B200Mi *mi;
  uint16_t *eob[3];
  int16_t *coef[3];
  B200BlockDec *dec[B200_LEVELS];     /* analysis decisions per level (0: 4x4, 1: 8x8, 2: 16x16, 3: 32x32) */
  int tr0, tr1, tc0, tc1;   /* current tile bounds in mi units */
} enc_t;

static const uint16_t *scan_for(int n) {
  return n == 4 ? av1_default_scan_4x4 : n == 8 ? av1_default_scan_8x8 : n == 16 ? av1_default_scan_16x16 : av1_default_scan_32x32;
}
static int txs_of(int n) { return n == 4 ? 0 : n == 8 ? 1 : n == 16 ? 2 : 3; }
static int bsize_of(int n) { return n == 4 ? BLOCK_4X4 : n == 8 ? BLOCK_8X8 : n == 16 ? BLOCK_16X16 : n == 32 ? BLOCK_32X32 : BLOCK_64X64; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

/* ---- intra edge preparation: AV1 spec 7.11.2 (no intra edge filter / upsampling: the sequence header sets
 * enable_intra_edge_filter = 0). above[-1..2n-1], left[-1..2n-1] via offset 16.  `r` is the plane the neighbours
 * are read from: the SOURCE plane in the open-loop analysis, the reconstruction in the final pass.  Top-right /
 * bottom-left availability is the static coding-order predicate b200_unit_before(). */
#define EOFF 16
static void build_edges(const enc_t *e, const uint16_t *r, int x, int y, int n, uint16_t *above, uint16_t *left,
                        int *have_above, int *have_left) {
  const B200Frame *f = e->f;
  const int st = e->stride;
  const int mi_r = y >> 2, mi_c = x >> 2, n4 = n >> 2;
  const int ha = mi_r > e->tr0, hl = mi_c > e->tc0;
  const int max_x = f->mi_cols * 4 - 1, max_y = f->mi_rows * 4 - 1;
  const int base = 1 << (f->bit_depth - 1);
  const int have_ar = ha && b200_unit_before(mi_r - 1, mi_c + n4, mi_r, mi_c, e->tr0, e->tr1, e->tc0, e->tc1);
  const int have_bl = hl && b200_unit_before(mi_r + n4, mi_c - 1, mi_r, mi_c, e->tr0, e->tr1, e->tc0, e->tc1);
  uint16_t *A = above + EOFF, *L = left + EOFF;
  if (!ha && hl) { for (int i = 0; i < 2 * n; i++) A[i] = r[y * st + x - 1]; }
  else if (!ha && !hl) { for (int i = 0; i < 2 * n; i++) A[i] = (uint16_t)(base - 1); }
  else {
    for (int i = 0; i < n; i++) { int xx = x + i; if (xx > max_x) xx = max_x; A[i] = r[(y - 1) * st + xx]; }
    for (int i = n; i < 2 * n; i++) {
      int xx = have_ar ? x + i : x + n - 1;
      if (xx > max_x) xx = max_x;
      A[i] = r[(y - 1) * st + xx];
    }
  }
  if (!hl && ha) { for (int i = 0; i < 2 * n; i++) L[i] = r[(y - 1) * st + x]; }
  else if (!hl && !ha) { for (int i = 0; i < 2 * n; i++) L[i] = (uint16_t)(base + 1); }
  else {
    for (int i = 0; i < n; i++) { int yy = y + i; if (yy > max_y) yy = max_y; L[i] = r[yy * st + x - 1]; }
    for (int i = n; i < 2 * n; i++) {
      int yy = have_bl ? y + i : y + n - 1;
      if (yy > max_y) yy = max_y;
      L[i] = r[yy * st + x - 1];
    }
  }
  uint16_t tl;
  if (ha && hl) tl = r[(y - 1) * st + x - 1];
  else if (ha) tl = r[(y - 1) * st + x];
  else if (hl) tl = r[y * st + x - 1];
  else tl = (uint16_t)base;
  A[-1] = tl; L[-1] = tl;
  *have_above = ha; *have_left = hl;
}

static const uint8_t k_bw4[22] = {1, 1, 2, 2, 2, 4, 4, 4, 8, 8, 8, 16, 16, 16, 32, 32, 1, 4, 2, 8, 4, 16};
static const uint8_t *sm_weights(int n) {
  return n == 4 ? av1_sm_weights_4 : n == 8 ? av1_sm_weights_8 : n == 16 ? av1_sm_weights_16 : n == 32 ? av1_sm_weights_32 : av1_sm_weights_64;
}

/* AV1 spec 7.11.2.x intra predictors for an n x n block. angle_delta in -3..3. */
/* ---- intra edge filter and upsampling of the directional predictors (AV1 spec 7.11.2.7 - 7.11.2.12), applied on
 * private copies of the edges.  filter_type: 0 / 1 = get_filter_type() (a smooth-mode neighbour above or left). */
static int ief_strength(int blk_wh, int filter_type, int delta) {                 /* 7.11.2.9 */
  const int d = abs(delta);
  int s = 0;
  if (filter_type == 0) {
    if (blk_wh <= 8) { if (d >= 56) s = 1; }
    else if (blk_wh <= 16) { if (d >= 40) s = 1; }
    else if (blk_wh <= 24) { if (d >= 8) s = 1; if (d >= 16) s = 2; if (d >= 32) s = 3; }
    else if (blk_wh <= 32) { if (d >= 1) s = 1; if (d >= 4) s = 2; if (d >= 32) s = 3; }
    else { if (d >= 1) s = 3; }
  } else {
    if (blk_wh <= 8) { if (d >= 40) s = 1; if (d >= 64) s = 2; }
    else if (blk_wh <= 16) { if (d >= 20) s = 1; if (d >= 48) s = 2; }
    else if (blk_wh <= 24) { if (d >= 4) s = 3; }
    else { if (d >= 1) s = 3; }
  }
  return s;
}
static int ief_upsample(int blk_wh, int filter_type, int delta) {                 /* 7.11.2.10 */
  const int d = abs(delta);
  if (d <= 0 || d >= 40) return 0;
  return filter_type == 0 ? blk_wh <= 16 : blk_wh <= 8;
}
static void ief_filter(uint16_t *buf /* index -1 .. sz-2 */, int sz, int strength) {   /* 7.11.2.12 */
  static const int kern[3][5] = {{0, 4, 8, 4, 0}, {0, 5, 6, 5, 0}, {2, 4, 4, 4, 2}};
  if (!strength) return;
  uint16_t edge[160];
  for (int i = 0; i < sz; i++) edge[i] = buf[i - 1];
  for (int i = 1; i < sz; i++) {
    int sum = 0;
    for (int j = 0; j < 5; j++) {
      int k = i - 2 + j;
      k = k < 0 ? 0 : (k > sz - 1 ? sz - 1 : k);
      sum += kern[strength - 1][j] * edge[k];
    }
    buf[i - 1] = (uint16_t)((sum + 8) >> 4);
  }
}
static void ief_upsample_edge(uint16_t *buf /* index -2 .. 2*num_px-2 */, int num_px, int bd) {   /* 7.11.2.11 */
  int dup[160];
  dup[0] = buf[-1];
  for (int i = -1; i < num_px; i++) dup[i + 2] = buf[i];
  dup[num_px + 2] = buf[num_px - 1];
  buf[-2] = (uint16_t)dup[0];
  const int mx = (1 << bd) - 1;
  for (int i = 0; i < num_px; i++) {
    int v = -dup[i] + 9 * dup[i + 1] + 9 * dup[i + 2] - dup[i + 3];
    v = (v + 8) >> 4;
    buf[2 * i - 1] = (uint16_t)(v < 0 ? 0 : (v > mx ? mx : v));
    buf[2 * i] = (uint16_t)dup[i + 2];
  }
}
/* directional prediction with the edge filter enabled (spec 7.11.2.4) */
static void predict_directional_ief(int p_angle, int n, const uint16_t *A0, const uint16_t *L0, int have_above, int have_left,
                                    int bit_depth, int filter_type, uint16_t *pred) {
  uint16_t Ab[EOFF + 160], Lb[EOFF + 160];
  uint16_t *A = Ab + EOFF, *L = Lb + EOFF;
  for (int i = -1; i < 2 * n; i++) { A[i] = A0[i]; L[i] = L0[i]; }
  int up_a = 0, up_l = 0;
  if (p_angle != 90 && p_angle != 180) {
    if (p_angle > 90 && p_angle < 180 && 2 * n >= 24) {          /* 7.11.2.7 corner filter */
      const int c = (L[0] * 5 + A[-1] * 6 + A[0] * 5 + 8) >> 4;
      A[-1] = L[-1] = (uint16_t)c;
    }
    if (have_above) ief_filter(A, n + (p_angle < 90 ? n : 0) + 1, ief_strength(2 * n, filter_type, p_angle - 90));
    if (have_left) ief_filter(L, n + (p_angle > 180 ? n : 0) + 1, ief_strength(2 * n, filter_type, p_angle - 180));
  }
  up_a = ief_upsample(2 * n, filter_type, p_angle - 90);
  if (up_a) ief_upsample_edge(A, n + (p_angle < 90 ? n : 0), bit_depth);
  up_l = ief_upsample(2 * n, filter_type, p_angle - 180);
  if (up_l) ief_upsample_edge(L, n + (p_angle > 180 ? n : 0), bit_depth);
  if (p_angle == 90) { for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) pred[i * n + j] = A[j]; return; }
  if (p_angle == 180) { for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) pred[i * n + j] = L[i]; return; }
  if (p_angle < 90) {
    const int dx = av1_dr_intra_derivative[p_angle], max_base = (2 * n - 1) << up_a;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const int idx = (i + 1) * dx, base = (idx >> (6 - up_a)) + (j << up_a), shift = ((idx << up_a) >> 1) & 31;
        pred[i * n + j] = base < max_base ? (uint16_t)((A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5) : A[max_base];
      }
  } else if (p_angle < 180) {
    const int dx = av1_dr_intra_derivative[180 - p_angle], dy = av1_dr_intra_derivative[p_angle - 90];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        int idx = (j << 6) - (i + 1) * dx, base = idx >> (6 - up_a);
        if (base >= -(1 << up_a)) {
          const int shift = ((idx << up_a) >> 1) & 31;
          pred[i * n + j] = (uint16_t)((A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5);
        } else {
          idx = (i << 6) - (j + 1) * dy; base = idx >> (6 - up_l);
          const int shift = ((idx << up_l) >> 1) & 31;
          pred[i * n + j] = (uint16_t)((L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5);
        }
      }
  } else {
    const int dy = av1_dr_intra_derivative[270 - p_angle];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const int idx = (j + 1) * dy, base = (idx >> (6 - up_l)) + (i << up_l), shift = ((idx << up_l) >> 1) & 31;
        pred[i * n + j] = (uint16_t)((L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5);
      }
  }
}

void orc_predict_intra_ef(int mode, int angle_delta, int n, const uint16_t *above, const uint16_t *left, int have_above,
                          int have_left, int bit_depth, int filter_type, uint16_t *pred);
/* edge filter disabled (sequence header enable_intra_edge_filter = 0; also the SATD pre-ranking of every frame) */
void orc_predict_intra(int mode, int angle_delta, int n, const uint16_t *above, const uint16_t *left, int have_above,
                       int have_left, int bit_depth, uint16_t *pred /* n*n */) {
  orc_predict_intra_ef(mode, angle_delta, n, above, left, have_above, have_left, bit_depth, -1, pred);
}
/* filter_type: -1 = edge filter off, 0 / 1 = on with that get_filter_type() result */
void orc_predict_intra_ef(int mode, int angle_delta, int n, const uint16_t *above, const uint16_t *left, int have_above,
                          int have_left, int bit_depth, int filter_type, uint16_t *pred /* n*n */) {
  const uint16_t *A = above + EOFF, *L = left + EOFF;
  if (filter_type >= 0 && mode >= V_PRED && mode <= D67_PRED) {
    predict_directional_ief(av1_mode_to_angle[mode] + angle_delta * 3, n, A, L, have_above, have_left, bit_depth, filter_type, pred);
    return;
  }
  if (mode == DC_PRED) {
    int v;
    if (have_above && have_left) { int s = 0; for (int i = 0; i < n; i++) s += A[i] + L[i]; v = (s + n) / (2 * n); }
    else if (have_left) { int s = 0; for (int i = 0; i < n; i++) s += L[i]; v = (s + (n >> 1)) / n; }
    else if (have_above) { int s = 0; for (int i = 0; i < n; i++) s += A[i]; v = (s + (n >> 1)) / n; }
    else v = 1 << (bit_depth - 1);
    for (int i = 0; i < n * n; i++) pred[i] = (uint16_t)v;
    return;
  }
  if (mode == PAETH_PRED) {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const int b = A[j] + L[i] - A[-1];
        const int pl = abs(b - L[i]), pt = abs(b - A[j]), ptl = abs(b - A[-1]);
        pred[i * n + j] = (pl <= pt && pl <= ptl) ? L[i] : (pt <= ptl ? A[j] : A[-1]);
      }
    return;
  }
  if (mode == SMOOTH_PRED || mode == SMOOTH_V_PRED || mode == SMOOTH_H_PRED) {
    const uint8_t *w = sm_weights(n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        int v;
        if (mode == SMOOTH_PRED)
          v = (w[i] * A[j] + (256 - w[i]) * L[n - 1] + w[j] * L[i] + (256 - w[j]) * A[n - 1] + 256) >> 9;
        else if (mode == SMOOTH_V_PRED) v = (w[i] * A[j] + (256 - w[i]) * L[n - 1] + 128) >> 8;
        else v = (w[j] * L[i] + (256 - w[j]) * A[n - 1] + 128) >> 8;
        pred[i * n + j] = (uint16_t)v;
      }
    return;
  }
  /* directional */
  const int p_angle = av1_mode_to_angle[mode] + angle_delta * 3;
  if (p_angle == 90) { for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) pred[i * n + j] = A[j]; return; }
  if (p_angle == 180) { for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) pred[i * n + j] = L[i]; return; }
  if (p_angle < 90) {
    const int dx = av1_dr_intra_derivative[p_angle];
    const int max_base = 2 * n - 1;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const int idx = (i + 1) * dx, base = (idx >> 6) + j, shift = (idx >> 1) & 31;
        pred[i * n + j] = base < max_base ? (uint16_t)((A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5) : A[max_base];
      }
  } else if (p_angle < 180) {
    const int dx = av1_dr_intra_derivative[180 - p_angle], dy = av1_dr_intra_derivative[p_angle - 90];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        int idx = (j << 6) - (i + 1) * dx, base = idx >> 6;
        if (base >= -1) {
          const int shift = (idx >> 1) & 31;
          pred[i * n + j] = (uint16_t)((A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5);
        } else {
          idx = (i << 6) - (j + 1) * dy; base = idx >> 6;
          const int shift = (idx >> 1) & 31;
          pred[i * n + j] = (uint16_t)((L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5);
        }
      }
  } else {
    const int dy = av1_dr_intra_derivative[270 - p_angle];
    const int max_base = 2 * n - 1;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const int idx = (j + 1) * dy, base = (idx >> 6) + i, shift = (idx >> 1) & 31;
        pred[i * n + j] = base < max_base ? (uint16_t)((L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5) : L[max_base];
      }
  }
}

/* sum of absolute 4x4 Hadamard coefficients over all 4x4 sub-blocks of an n x n residual */
static uint32_t satd4(const uint16_t *src, int sstride, const uint16_t *pred, int n) {
  uint32_t total = 0;
  for (int by = 0; by < n; by += 4)
    for (int bx = 0; bx < n; bx += 4) {
      int d[16], t[16];
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) d[i * 4 + j] = (int)src[(by + i) * sstride + bx + j] - (int)pred[(by + i) * n + bx + j];
      for (int i = 0; i < 4; i++) {
        const int a = d[i * 4] + d[i * 4 + 1], b = d[i * 4] - d[i * 4 + 1], c = d[i * 4 + 2] + d[i * 4 + 3], e = d[i * 4 + 2] - d[i * 4 + 3];
        t[i * 4] = a + c; t[i * 4 + 1] = b + e; t[i * 4 + 2] = a - c; t[i * 4 + 3] = b - e;
      }
      for (int j = 0; j < 4; j++) {
        const int a = t[j] + t[4 + j], b = t[j] - t[4 + j], c = t[8 + j] + t[12 + j], e = t[8 + j] - t[12 + j];
        total += (uint32_t)(abs(a + c) + abs(b + e) + abs(a - c) + abs(b - e));
      }
    }
  return total;
}

/* quantise -> levels in scan order + eob; dequantise -> dq (row-major).  Levels clamped to int16. */
/* coefficient position (row-major) of scan index i: default zig-zag-like scan for the 2-D transform class, row by row
 * (mrow) for TX_CLASS_VERT (V_DCT), column by column (mcol) for TX_CLASS_HORIZ (H_DCT): AV1 spec get_scan() */
static inline int scan_pos(const uint16_t *scan, int n, int tx_class, int i) {
  return tx_class == 0 ? scan[i] : (tx_class == 2 ? i : (i % n) * n + i / n);
}
static int quantize(const B200Frame *f, const int32_t *coef, int n, int tx_type, int16_t *qc_scan, int32_t *dq) {
  const uint16_t *scan = scan_for(n);
  const int n2 = n * n, txc = b200_tx_class(tx_type);
  int eob = 0;
  const int shift = n == 32 ? 1 : 0;   /* 32x32: forward scale is half, dequant >> 1 */
  for (int i = 0; i < n2; i++) {
    const int pos = scan_pos(scan, n, txc, i);
    const int q = pos == 0 ? f->dc_q : f->ac_q;
    const int off = pos == 0 ? (q * 109) >> 8 : (q * 98) >> 8;
    const int32_t c = coef[pos];
    const uint32_t a = (uint32_t)(c < 0 ? -c : c);
    uint32_t lv = (uint32_t)(((uint64_t)(a << shift) + (uint32_t)off) / (uint32_t)q);
    if (lv > 32767) lv = 32767;
    const int32_t dqa = (int32_t)(((uint64_t)lv * (uint32_t)q) & 0xFFFFFF) >> shift;
    qc_scan[i] = (int16_t)(c < 0 ? -(int)lv : (int)lv);
    dq[pos] = c < 0 ? -dqa : dqa;
    if (lv) eob = i + 1;
  }
  return eob;
}

/* static-CDF rate of one transform block (mirrors the packer's symbol sequence), 1/256 bit */
static uint32_t coef_rate(const enc_t *e, int plane, int n, int tx_type, int y_mode, const int16_t *qc, int eob) {
  const B200RateTables *rt = e->rt;
  const int txs = txs_of(n), ptype = plane > 0;
  const int skip_ctx = plane == 0 ? 0 : 7;
  if (eob == 0) return rt->txb_skip[txs][skip_ctx][1];
  uint32_t r = rt->txb_skip[txs][skip_ctx][0];
  if (plane == 0 && n <= 16) r += rt->intra_tx[(e->f->reduced_tx_set || n == 16) ? 1 : 0][txs][y_mode][tx_type];
  static const int16_t group_start[12] = {0, 1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 513};
  static const int8_t offset_bits[12] = {0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9};
  int t = 0;
  while (t < 11 && group_start[t + 1] <= eob) t++;
  const int txc = b200_tx_class(tx_type);
  r += rt->eob_pt[txs][ptype][txc != 0][t - 1];
  if (offset_bits[t] > 0) {
    const int extra = eob - group_start[t], nb = offset_bits[t];
    r += rt->eob_extra[txs][ptype][t - 3][(extra >> (nb - 1)) & 1];
    r += 256u * (uint32_t)(nb - 1);
  }
  const uint16_t *scan = scan_for(n);
  const int stride = n + 4;
  uint8_t lev[36 * 36];
  memset(lev, 0, (size_t)stride * (n + 4));
  for (int i = 0; i < eob; i++) {
    const int pos = scan_pos(scan, n, txc, i);
    const int a = abs(qc[i]);
    lev[(pos / n) * stride + pos % n] = (uint8_t)(a > 127 ? 127 : a);
  }
  for (int i = eob - 1; i >= 0; i--) {
    const int pos = scan_pos(scan, n, txc, i), rr = pos / n, cc = pos % n;
    const int level = abs(qc[i]);
    const uint8_t *L = lev + rr * stride + cc;
#define M3(v) ((v) > 3 ? 3 : (v))
#define M15(v) ((v) > 15 ? 15 : (v))
    if (i == eob - 1) {
      const int c4 = i == 0 ? 0 : (i <= (n * n) / 8 ? 1 : (i <= (n * n) / 4 ? 2 : 3));
      r += rt->coeff_base_eob[txs][ptype][c4][M3(level) - 1];
    } else {
      /* neighbourhood and offset by transform class (AV1 spec Sig_Ref_Diff_Offset / Coeff_Base_Pos_Ctx_Offset) */
      int mag = M3(L[1]) + M3(L[stride]);
      if (txc == 0) mag += M3(L[stride + 1]) + M3(L[2]) + M3(L[2 * stride]);
      else if (txc == 1) mag += M3(L[2]) + M3(L[3]) + M3(L[4]);
      else mag += M3(L[2 * stride]) + M3(L[3 * stride]) + M3(L[4 * stride]);
      int ctx = (mag + 1) >> 1;
      if (ctx > 4) ctx = 4;
      if (txc == 0) {
        if (rr == 0 && cc == 0) ctx = 0;
        else ctx += (rr + cc < 2) ? 1 : (rr + cc < 4) ? 6 : 21;
      } else {
        const int idx = txc == 2 ? rr : cc;
        ctx += 26 + 5 * (idx > 2 ? 2 : idx);
      }
      r += rt->coeff_base[txs][ptype][ctx][M3(level)];
    }
    if (level > 2) {
      int mag = M15(L[1]) + M15(L[stride]) + M15(txc == 0 ? L[stride + 1] : (txc == 1 ? L[2] : L[2 * stride]));
      mag = (mag + 1) >> 1;
      if (mag > 6) mag = 6;
      const int near = txc == 0 ? (rr < 2 && cc < 2) : (txc == 1 ? cc == 0 : rr == 0);
      const int bctx = (rr == 0 && cc == 0) ? mag : (near ? mag + 7 : mag + 14);
      const int br = level - 3;
      r += rt->coeff_br[txs > 3 ? 3 : txs][ptype][bctx][br > 12 ? 12 : br];
      if (level > 14) {   /* golomb(level - 15): 2*len - 1 bits */
        int x = level - 15 + 1, len = 0;
        while (x) { x >>= 1; len++; }
        r += 256u * (uint32_t)(2 * len - 1);
      }
    }
    if (level) r += (i == 0) ? rt->dc_sign[ptype][0][qc[i] < 0] : 256u;
  }
  return r;
}

static uint64_t sse_block(const uint16_t *src, int sstride, const uint16_t *rec, int n) {
  uint64_t s = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { const int d = (int)src[i * sstride + j] - (int)rec[i * n + j]; s += (uint64_t)(d * d); }
  return s;
}

/* full trial of one plane: pred -> residual -> fwd -> quant -> rate, dequant -> inverse -> recon, SSE.
 * recon (n*n) receives the reconstruction. returns distortion; *rate out. */
static uint64_t trial_plane(const enc_t *e, int plane, int x, int y, int n, const uint16_t *pred, int tx_type, int y_mode,
                            int16_t *qc, uint16_t *eob_out, uint16_t *recon, uint32_t *rate) {
  const B200Frame *f = e->f;
  const uint16_t *src = e->src[plane] + y * e->stride + x;
  int16_t resid[32 * 32];
  int32_t coef[32 * 32], dq[32 * 32];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) resid[i * n + j] = (int16_t)((int)src[i * e->stride + j] - (int)pred[i * n + j]);
  orc_fwd_txfm2d(resid, n, coef, n, tx_type);
  const int eob = quantize(f, coef, n, tx_type, qc, dq);
  memcpy(recon, pred, sizeof(uint16_t) * n * n);
  if (eob) orc_inv_txfm2d_add(dq, recon, n, n, tx_type, f->bit_depth);
  *eob_out = (uint16_t)eob;
  *rate = coef_rate(e, plane, n, tx_type, y_mode, qc, eob);
  if (f->psy && plane == 0 && n >= 8) {      /* Tune::Psychovisual: variance-aware error per 8x8 (records.h b200_psy_dist8x8) */
    uint64_t total = 0;
    for (int by = 0; by < n; by += 8)
      for (int bx = 0; bx < n; bx += 8) {
        uint32_t ss = 0, sd = 0, s2 = 0, d2 = 0, sp = 0;
        for (int i = 0; i < 8; i++)
          for (int j = 0; j < 8; j++) {
            const uint32_t s = src[(by + i) * e->stride + bx + j], d = recon[(by + i) * n + bx + j];
            ss += s; sd += d; s2 += s * s; d2 += d * d; sp += s * d;
          }
        total += b200_psy_dist8x8(ss, sd, s2, d2, sp, f->bit_depth);
      }
    return total > 0xffffffffu ? 0xffffffffu : total;
  }
  return sse_block(src, e->stride, recon, n);
}

static inline int64_t rd_cost(const B200Frame *f, uint64_t sse, uint32_t rate) {
  return (int64_t)(sse << 8) + (((int64_t)f->lambda_q8 * (int64_t)rate) >> 8);
}

static const uint8_t k_tx_types[B200_TX_TYPES] = {B200_TX_TYPE_LIST};

/* chroma transform type derived from the uv mode (AV1 spec get_tx_type for intra chroma),
 * restricted to what the (reduced) intra set allows for this size */
static int chroma_tx_type(int uv_mode, int n) {
  if (uv_mode == UV_CFL_PRED) return TX_DCT_DCT;
  if (n >= 32) return TX_DCT_DCT;
  if (uv_mode == 13) return TX_DCT_DCT;           /* UV_CFL_PRED */
  return av1_intra_mode_to_tx_type[uv_mode];   /* DCT/ADST combinations are in every intra set up to 16x16 */
}

static uint32_t mode_rate(const B200RateTables *rt, const uint16_t *row, int m, int n, int delta) {
  uint32_t r = row[m];
  if (n >= 8 && m >= V_PRED && m <= D67_PRED) r += rt->angle_delta[m - V_PRED][delta + 3];
  return r;
}

/* ---- chroma from luma (AV1 spec 7.11.5, 4:4:4 so no subsampling).  rav1e considers CFL for every intra chroma block
 * whose size allows it (rdo.rs rdo_cfl_alpha); alpha is the integer least-squares fit, clamped to +-16. */
static void cfl_ac(const uint16_t *luma, int stride, int n, int16_t *ac) {
  int sum = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) sum += luma[i * stride + j] << 3;
  const int avg = (sum + ((n * n) >> 1)) >> (2 * ilog2(n));
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) ac[i * n + j] = (int16_t)((luma[i * stride + j] << 3) - avg);
}
static int cfl_pick_alpha(const int16_t *ac, const uint16_t *src, int stride, int n, int dc) {
  int64_t num = 0, den = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      const int a = ac[i * n + j];
      num += (int64_t)a * ((int)src[i * stride + j] - dc);
      den += (int64_t)a * a;
    }
  if (den == 0) return 0;
  const uint64_t mag = (((uint64_t)(num < 0 ? -num : num) << 6) + (uint64_t)(den >> 1)) / (uint64_t)den;
  const int a = mag > 16 ? 16 : (int)mag;
  return num < 0 ? -a : a;
}
static void cfl_predict(const int16_t *ac, int n, int dc, int alpha, int bit_depth, uint16_t *pred) {
  const int mx = (1 << bit_depth) - 1;
  for (int i = 0; i < n * n; i++) {
    const int sc = alpha * ac[i];
    int v = dc + (sc < 0 ? -((-sc + 32) >> 6) : ((sc + 32) >> 6));
    pred[i] = (uint16_t)(v < 0 ? 0 : (v > mx ? mx : v));
  }
}
/* joint sign (0..7) and alpha indices of a signed alpha pair (not both zero) */
static void cfl_code(int au, int av, int *js, int *idx) {
  const int su = au == 0 ? 0 : (au < 0 ? 1 : 2), sv = av == 0 ? 0 : (av < 0 ? 1 : 2);
  *js = su * 3 + sv - 1;
  *idx = (au ? abs(au) - 1 : 0) | ((av ? abs(av) - 1 : 0) << 4);
}
static uint32_t cfl_rate(const B200RateTables *rt, int au, int av) {
  const int su = au == 0 ? 0 : (au < 0 ? 1 : 2), sv = av == 0 ? 0 : (av < 0 ? 1 : 2);
  uint32_t r = rt->cfl_sign[su * 3 + sv - 1];
  if (su) r += rt->cfl_alpha[(su - 1) * 3 + sv][abs(au) - 1];
  if (sv) r += rt->cfl_alpha[(sv - 1) * 3 + su][abs(av) - 1];
  return r;
}
static void cfl_decode(int js, int idx, int *au, int *av) {
  const int su = (js + 1) / 3, sv = (js + 1) % 3;
  *au = su == 0 ? 0 : (su == 1 ? -1 : 1) * ((idx & 15) + 1);
  *av = sv == 0 ? 0 : (sv == 1 ? -1 : 1) * ((idx >> 4) + 1);
}

/* fine_directional_intra (ravif: speed <= 6, av1encoder.rs:585): refine the angle of the best nominal directional
 * mode of a plane group by SATD.  sat[]: nominal SATDs (updated in place for the refined mode); returns the packed
 * (mode | (delta + 3) << 8) word, mode 255 = nothing refined.  Candidate order 0, -3, -2, -1, 1, 2, 3; strict '<'. */
static uint32_t refine_angle(enc_t *e, int first_plane, int last_plane, int x, int y, int n, uint16_t A[3][EOFF + 66],
                             uint16_t L[3][EOFF + 66], int ha, int hl, int mask, uint32_t sat[13]) {
  const B200Frame *f = e->f;
  if (!f->fine_directional_intra || n < 8) return 255u | (3u << 8);
  int md = -1;
  for (int m = V_PRED; m <= D67_PRED; m++)
    if (((mask >> m) & 1) && (md < 0 || sat[m] < sat[md])) md = m;
  if (md < 0) return 255u | (3u << 8);
  uint16_t pred[32 * 32];
  int best_d = 0;
  int64_t best = (int64_t)sat[md] * 64 + (((int64_t)f->lambda_satd_q8 * e->rt->angle_delta[md - V_PRED][3]) >> 8);
  uint32_t best_s = sat[md];
  static const int order[6] = {-3, -2, -1, 1, 2, 3};
  for (int k = 0; k < 6; k++) {
    const int d = order[k];
    uint32_t sv = 0;
    for (int p = first_plane; p <= last_plane; p++) {
      orc_predict_intra(md, d, n, A[p], L[p], ha, hl, f->bit_depth, pred);
      sv += satd4(e->src[p] + y * e->stride + x, e->stride, pred, n);
    }
    const int64_t c = (int64_t)sv * 64 + (((int64_t)f->lambda_satd_q8 * e->rt->angle_delta[md - V_PRED][d + 3]) >> 8);
    if (c < best) { best = c; best_d = d; best_s = sv; }
  }
  sat[md] = best_s;
  return (uint32_t)md | ((uint32_t)(best_d + 3) << 8);
}

/* ---- stage A: open-loop analysis of one block.  Every prediction reads SOURCE neighbours, so all blocks of all
 * levels are independent of each other (one massively parallel kernel per level on the device).  Rate terms that
 * would depend on neighbouring decisions use the context-0 tables. */
void orc_analyze_block(enc_t *e, int x, int y, int n, B200BlockDec *out, uint32_t satd_y[14]) {
  const B200Frame *f = e->f;
  const B200RateTables *rt = e->rt;
  uint16_t A[3][EOFF + 66], L[3][EOFF + 66], pred[32 * 32];
  int ha = 0, hl = 0;
  for (int p = 0; p < f->planes; p++) build_edges(e, e->src[p], x, y, n, A[p], L[p], &ha, &hl);
  const uint16_t *ymode_cost = rt->y_mode[0][0];
  /* open loop: the neighbours' modes are not known yet, the trials assume get_filter_type() = 0; the SATD pre-ranking
   * and the angle refinement use the unfiltered edges */
  const int ief = f->enable_intra_edge_filter ? 0 : -1;
  /* luma SATD of the nominal modes (the device computes this table in the SATD kernels) */
  for (int m = 0; m < 13; m++) {
    satd_y[m] = 0;
    if (!((f->y_mode_mask >> m) & 1)) continue;
    orc_predict_intra(m, 0, n, A[0], L[0], ha, hl, f->bit_depth, pred);
    satd_y[m] = satd4(e->src[0] + y * e->stride + x, e->stride, pred, n);
  }
  satd_y[13] = refine_angle(e, 0, 0, x, y, n, A, L, ha, hl, f->y_mode_mask, satd_y);
  const int md_y = satd_y[13] & 255, dl_y = (int)(satd_y[13] >> 8) - 3;
#define DELTA_Y(m) ((m) == md_y ? dl_y : 0)
  /* chroma follows rav1e's candidate structure (rdo.rs rdo_mode_decision): the chosen luma mode (with its angle
   * delta), DC_PRED, and chroma-from-luma - no chroma SATD ranking */
#define DELTA_UV(m) ((m) == best_mode ? DELTA_Y(m) : 0)
  /* ranking: SATD + mode rate */
  /* the K = frame.luma_rdo_modes (2 or 4) best modes, lowest cost first, ties -> lowest mode index */
  const int K = f->luma_rdo_modes >= 4 ? 4 : 2;
  int64_t bc[4] = {INT64_MAX, INT64_MAX, INT64_MAX, INT64_MAX};
  int bm[4] = {-1, -1, -1, -1};
  for (int m = 0; m < 13; m++) {
    if (!((f->y_mode_mask >> m) & 1)) continue;
    const int64_t c = (int64_t)satd_y[m] * 64 + (((int64_t)f->lambda_satd_q8 * mode_rate(rt, ymode_cost, m, n, DELTA_Y(m))) >> 8);
    int at = K;
    while (at > 0 && c < bc[at - 1]) at--;
    for (int j = K - 1; j > at; j--) { bc[j] = bc[j - 1]; bm[j] = bm[j - 1]; }
    if (at < K) { bc[at] = c; bm[at] = m; }
  }
  /* luma trials: the ranked modes, pair by pair (ranks 0-1, then 2-3), x allowed transform types; order (mode rank, type
   * index), strict '<' keeps the earliest */
  /* 32x32: the intra transform set is DCT only (AV1 spec get_tx_set: no tx_type is coded) */
  const int tx_mask = b200_tx_mask_for(f->tx_type_mask, f->reduced_tx_set, n);
  int first_tt = -1;
  for (int k = 0; k < B200_TX_TYPES; k++) if ((tx_mask >> k) & 1) { first_tt = k_tx_types[k]; break; }
  int64_t bestJ = INT64_MAX;
  int best_mode = bm[0], best_tt = first_tt, eob_y = 0, best_c = 1 << 20;
  int16_t qc[1024];
  uint16_t recon[1024], eobv;
  /* pass 0: the first allowed transform type on both ranked modes; pass 1: the other types - skipped when every
   * first-type trial quantised to all-zero (nothing left for another transform to gain).  Winner = lowest J,
   * ties -> lowest candidate index c = rank * B200_TX_TYPES + type. */
  for (int pr = 0; pr < K / 2; pr++) {
  int all_first_zero = 1;       /* per pair of ranked modes */
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1 && (all_first_zero || !(f->rdo_tx_decision && n <= 16))) break;
    for (int k = 2 * pr; k < 2 * pr + 2; k++) {
      const int m = bm[k];
      if (m < 0) continue;
      orc_predict_intra_ef(m, DELTA_Y(m), n, A[0], L[0], ha, hl, f->bit_depth, ief, pred);
      const uint32_t mrate = mode_rate(rt, ymode_cost, m, n, DELTA_Y(m));
      for (int t = 0; t < B200_TX_TYPES; t++) {
        const int tt = k_tx_types[t];
        if (!((tx_mask >> t) & 1)) continue;
        if ((pass == 0) != (tt == first_tt)) continue;
        uint32_t rate;
        const uint64_t d = trial_plane(e, 0, x, y, n, pred, tt, m, qc, &eobv, recon, &rate);
        if (pass == 0 && eobv) all_first_zero = 0;
        const int64_t J = rd_cost(f, d, rate + mrate);
        const int c = k * B200_TX_TYPES + t;
        if (J < bestJ || (J == bestJ && c < best_c)) { bestJ = J; best_c = c; best_mode = m; best_tt = tt; eob_y = eobv; }
      }
    }
  }
  }
  int64_t total = bestJ;
  int uv_best = DC_PRED, eob_u = 0, eob_v = 0, cfl_js = 0, cfl_idx = 0;
  if (f->planes > 1) {
    const uint16_t *uvcost = rt->uv_mode[1][best_mode];
    int um[2] = {-1, -1};
    if ((f->uv_mode_mask >> best_mode) & 1) um[0] = best_mode;
    if (best_mode != DC_PRED && (f->uv_mode_mask & 1)) { if (um[0] < 0) um[0] = DC_PRED; else um[1] = DC_PRED; }
    if (um[0] < 0)      /* neither is allowed (conformance tests force single modes): the lowest allowed mode */
      for (int m = 0; m < 13; m++) if ((f->uv_mode_mask >> m) & 1) { um[0] = m; break; }
    int64_t bestC = INT64_MAX;
    for (int k = 0; k < 2; k++) {
      const int m = um[k];
      if (m < 0) continue;
      const int ctt = chroma_tx_type(m, n);
      uint16_t e1, e2;
      uint32_t r1, r2;
      orc_predict_intra_ef(m, DELTA_UV(m), n, A[1], L[1], ha, hl, f->bit_depth, ief, pred);
      uint64_t d = trial_plane(e, 1, x, y, n, pred, ctt, best_mode, qc, &e1, recon, &r1);
      orc_predict_intra_ef(m, DELTA_UV(m), n, A[2], L[2], ha, hl, f->bit_depth, ief, pred);
      d += trial_plane(e, 2, x, y, n, pred, ctt, best_mode, qc, &e2, recon, &r2);
      const int64_t J = rd_cost(f, d, r1 + r2 + mode_rate(rt, uvcost, m, n, DELTA_UV(m)));
      if (J < bestC) { bestC = J; uv_best = m; eob_u = e1; eob_v = e2; }
    }
    if (f->enable_cfl) {      /* third candidate: chroma from (source) luma */
      int16_t ac[1024];
      int al[2], dcv[2];
      cfl_ac(e->src[0] + y * e->stride + x, e->stride, n, ac);
      for (int p = 1; p < 3; p++) {
        orc_predict_intra(DC_PRED, 0, n, A[p], L[p], ha, hl, f->bit_depth, pred);
        dcv[p - 1] = pred[0];
        al[p - 1] = cfl_pick_alpha(ac, e->src[p] + y * e->stride + x, e->stride, n, dcv[p - 1]);
      }
      if (al[0] || al[1]) {
        uint16_t e1, e2;
        uint32_t r1, r2;
        cfl_predict(ac, n, dcv[0], al[0], f->bit_depth, pred);
        uint64_t d = trial_plane(e, 1, x, y, n, pred, TX_DCT_DCT, best_mode, qc, &e1, recon, &r1);
        cfl_predict(ac, n, dcv[1], al[1], f->bit_depth, pred);
        d += trial_plane(e, 2, x, y, n, pred, TX_DCT_DCT, best_mode, qc, &e2, recon, &r2);
        const int64_t J = rd_cost(f, d, r1 + r2 + uvcost[UV_CFL_PRED] + cfl_rate(rt, al[0], al[1]));
        if (J < bestC) { bestC = J; uv_best = UV_CFL_PRED; eob_u = e1; eob_v = e2; cfl_code(al[0], al[1], &cfl_js, &cfl_idx); }
      }
    }
    total += bestC;
  }
  const int all_zero = eob_y == 0 && (f->planes == 1 || (eob_u == 0 && eob_v == 0));
  total += ((int64_t)f->lambda_q8 * rt->skip[0][all_zero]) >> 8;
  memset(out, 0, sizeof(*out));
  out->cost = total;
  out->y_mode = (uint8_t)best_mode; out->tx_type = (uint8_t)best_tt; out->uv_mode = (uint8_t)uv_best;
  out->all_zero = (uint8_t)all_zero;
  out->angle_y = (uint8_t)(DELTA_Y(best_mode) + 3);
  out->angle_uv = (uint8_t)(DELTA_UV(uv_best) + 3);
  out->cfl_js = (uint8_t)cfl_js; out->cfl_alpha = (uint8_t)cfl_idx;
#undef DELTA_Y
#undef DELTA_UV
}

/* ---- stage B: partition decision of one 16x16 region from the analysed costs (context-0 partition rates) */
static int64_t part_cost(const B200Frame *f, const B200RateTables *rt, int mi_r, int mi_c, int n, int split) {
  const int bsl = ilog2(n) - 2, hbs = n >> 3;
  if (!((mi_r + hbs) < f->mi_rows && (mi_c + hbs) < f->mi_cols)) return 0;   /* forced / 1-bit cases ignored */
  return ((int64_t)f->lambda_q8 * rt->partition[(bsl - 1) * 4][split ? 3 : 0]) >> 8;
}
static void set_mi(enc_t *e, int x, int y, int n, const B200BlockDec *d) {
  const B200Frame *f = e->f;
  B200Mi m;
  memset(&m, 0, sizeof(m));
  m.bsize = (uint8_t)bsize_of(n); m.y_mode = d->y_mode; m.uv_mode = d->uv_mode;
  m.angle = (uint8_t)((d->angle_y ? d->angle_y : 3) | ((d->angle_uv ? d->angle_uv : 3) << 4));
  m.tx_size = (uint8_t)txs_of(n); m.tx_type = d->tx_type;
  if (d->uv_mode == UV_CFL_PRED) { m.flags = (uint8_t)(d->cfl_js << 1); m.cfl = d->cfl_alpha; }
  for (int i = 0; i < n / 4; i++)
    for (int j = 0; j < n / 4; j++) e->mi[((y >> 2) + i) * f->mi_cols + (x >> 2) + j] = m;
}
static const B200BlockDec *dec_at(const enc_t *e, int lv, int x, int y) {
  const int n = 4 << lv;
  return &e->dec[lv][(y / n) * b200_level_cols(e->f, lv) + (x / n)];
}
/* Best RD cost of the block (level lv, top-left x, y): PARTITION_NONE against PARTITION_SPLIT into four blocks of the
 * level below, recursively; blocks that stick out of the (8-pixel aligned) frame must split.  commit != 0 also writes
 * the winning decisions into the mode-info array.  Ties keep the larger block. */
static int64_t decide_block(enc_t *e, int lv, int x, int y, int commit) {
  const B200Frame *f = e->f;
  const int W = f->mi_cols * 4, H = f->mi_rows * 4, n = 4 << lv;
  if (x >= W || y >= H) return 0;
  const int inside = (x + n <= W) && (y + n <= H);
  const int can_none = inside && n <= f->partition_max;
  const int can_split = lv > 0 && (f->partition_min < n || !inside);
  int64_t j_none = 0, j_split = 0;
  if (can_none) j_none = dec_at(e, lv, x, y)->cost + (lv > 0 ? part_cost(f, e->rt, y >> 2, x >> 2, n, 0) : 0);
  if (can_split) {
    j_split = part_cost(f, e->rt, y >> 2, x >> 2, n, 1);
    for (int k = 0; k < 4; k++) j_split += decide_block(e, lv - 1, x + (k & 1) * (n >> 1), y + (k >> 1) * (n >> 1), 0);
  }
  const int none = can_none && (!can_split || j_none <= j_split);
  if (commit) {
    if (none) set_mi(e, x, y, n, dec_at(e, lv, x, y));
    else for (int k = 0; k < 4; k++) decide_block(e, lv - 1, x + (k & 1) * (n >> 1), y + (k >> 1) * (n >> 1), 1);
  }
  return none ? j_none : j_split;
}
/* root regions: 32x32 when the partition range reaches it, else 16x16 */
static int root_level(const B200Frame *f) { return f->partition_max >= 32 ? 3 : 2; }
void orc_decide_region(enc_t *e, int x, int y) { decide_block(e, root_level(e->f), x, y, 1); }

/* ---- stage C: closed-loop reconstruction of one decided block (the only stage that needs reconstructed neighbours) */
static void reconstruct_block(enc_t *e, int x, int y, int n) {
  const B200Frame *f = e->f;
  const int mi_r = y >> 2, mi_c = x >> 2, n4 = n >> 2;
  B200Mi m = e->mi[mi_r * f->mi_cols + mi_c];
  uint16_t A[EOFF + 66], L[EOFF + 66], pred[32 * 32], recon[32 * 32];
  int16_t qc[1024];
  int all_zero = 1;
  for (int p = 0; p < f->planes; p++) {
    int ha, hl;
    build_edges(e, e->rec[p], x, y, n, A, L, &ha, &hl);
    const int mode = p == 0 ? m.y_mode : m.uv_mode;
    const int tt = p == 0 ? m.tx_type : chroma_tx_type(m.uv_mode, n);
    if (p > 0 && mode == UV_CFL_PRED) {      /* DC + alpha * AC of the RECONSTRUCTED luma of this block */
      int16_t ac[1024];
      int au, av;
      cfl_decode((m.flags >> 1) & 7, m.cfl, &au, &av);
      cfl_ac(e->rec[0] + y * e->stride + x, e->stride, n, ac);
      orc_predict_intra(DC_PRED, 0, n, A, L, ha, hl, f->bit_depth, pred);
      cfl_predict(ac, n, pred[0], p == 1 ? au : av, f->bit_depth, pred);
    } else {
      int ief = -1;
      if (f->enable_intra_edge_filter) {      /* get_filter_type (spec 7.11.2.8): a smooth mode above or to the left */
        ief = 0;
        for (int side = 0; side < 2; side++) {
          if (side == 0 ? !ha : !hl) continue;
          const B200Mi *nb = &e->mi[(side == 0 ? mi_r - 1 : mi_r) * f->mi_cols + (side == 0 ? mi_c : mi_c - 1)];
          const int nm = p == 0 ? nb->y_mode : nb->uv_mode;
          if (nm == SMOOTH_PRED || nm == SMOOTH_V_PRED || nm == SMOOTH_H_PRED) ief = 1;
        }
      }
      orc_predict_intra_ef(mode, (p == 0 ? (m.angle & 15) : (m.angle >> 4)) - 3, n, A, L, ha, hl, f->bit_depth, ief, pred);
    }
    uint16_t eobv;
    uint32_t rate;
    trial_plane(e, p, x, y, n, pred, tt, m.y_mode, qc, &eobv, recon, &rate);
    for (int i = 0; i < n; i++) memcpy(e->rec[p] + (y + i) * e->stride + x, recon + i * n, sizeof(uint16_t) * n);
    e->eob[p][mi_r * f->mi_cols + mi_c] = eobv;
    memcpy(e->coef[p] + b200_coef_offset(f, mi_r, mi_c), qc, sizeof(int16_t) * n * n);
    if (eobv) all_zero = 0;
  }
  if (all_zero) m.flags |= 1;
  for (int i = 0; i < n4; i++)
    for (int j = 0; j < n4; j++) e->mi[(mi_r + i) * f->mi_cols + mi_c + j] = m;
}

static void tile_bounds(enc_t *e, int tile_row, int tile_col) {
  const B200Frame *f = e->f;
  e->tr0 = f->tile_row_start_sb[tile_row] * 16;
  e->tr1 = f->tile_row_start_sb[tile_row + 1] * 16; if (e->tr1 > f->mi_rows) e->tr1 = f->mi_rows;
  e->tc0 = f->tile_col_start_sb[tile_col] * 16;
  e->tc1 = f->tile_col_start_sb[tile_col + 1] * 16; if (e->tc1 > f->mi_cols) e->tc1 = f->mi_cols;
}
static void tile_of(const B200Frame *f, int mi_r, int mi_c, int *tr, int *tc) {
  int r = 0, c = 0;
  while (f->tile_row_start_sb[r + 1] * 16 <= mi_r) r++;
  while (f->tile_col_start_sb[c + 1] * 16 <= mi_c) c++;
  *tr = r; *tc = c;
}

/* Whole frame: analysis of every block at every allowed level, partition decision, reconstruction chain.
 * dec[lv] (may be NULL): receives the analysis decisions, satd (may be NULL): [lv][block][2][13] for tests. */
void orc_encode_frame(const B200Frame *f, const B200RateTables *rt, uint16_t *const src[3], uint16_t *const rec[3],
                      B200Mi *mi, uint16_t *const eob[3], int16_t *const coef[3], B200BlockDec *const dec_out[B200_LEVELS],
                      uint32_t *const satd_out[B200_LEVELS]) {
  enc_t e;
  memset(&e, 0, sizeof(e));
  e.f = f; e.rt = rt; e.stride = f->mi_cols * 4; e.mi = mi;
  for (int p = 0; p < 3; p++) { e.src[p] = src[p]; e.rec[p] = rec[p]; e.eob[p] = eob[p]; e.coef[p] = coef[p]; }
  for (int lv = 0; lv < B200_LEVELS; lv++) {
    const size_t nblk = (size_t)b200_level_rows(f, lv) * b200_level_cols(f, lv);
    e.dec[lv] = (dec_out && dec_out[lv]) ? dec_out[lv] : (B200BlockDec *)calloc(nblk ? nblk : 1, sizeof(B200BlockDec));
  }
  for (int lv = B200_LEVELS - 1; lv >= 0; lv--) {     /* large to small: a level is pruned by its parent level */
    const int n = 4 << lv;
    /* 8x8 is always analysed: 16x16 regions that stick out of the frame are forced to split down to 8x8 */
    if (n < (f->partition_min < 8 ? f->partition_min : 8) || n > f->partition_max) continue;
    for (int by = 0; by < b200_level_rows(f, lv); by++)
      for (int bx = 0; bx < b200_level_cols(f, lv); bx++) {
        int tr, tc;
        tile_of(f, by * (n / 4), bx * (n / 4), &tr, &tc);
        tile_bounds(&e, tr, tc);
        uint32_t sy[14];
        const size_t bi = (size_t)by * b200_level_cols(f, lv) + bx;
        if (lv < B200_LEVELS - 1 && b200_pruned_by_parent(f, lv, bx, by, e.dec[lv + 1])) {
          memset(&e.dec[lv][bi], 0, sizeof(B200BlockDec));
          e.dec[lv][bi].cost = B200_COST_PRUNED;
          if (satd_out && satd_out[lv]) memset(satd_out[lv] + bi * B200_SATD_WORDS, 0, B200_SATD_WORDS * sizeof(uint32_t));
          continue;
        }
        orc_analyze_block(&e, bx * n, by * n, n, &e.dec[lv][bi], sy);
        if (satd_out && satd_out[lv]) memcpy(satd_out[lv] + bi * B200_SATD_WORDS, sy, sizeof(sy));
      }
  }
  const int rn = 4 << root_level(f);
  for (int y = 0; y < f->mi_rows * 4; y += rn)
    for (int x = 0; x < f->mi_cols * 4; x += rn) orc_decide_region(&e, x, y);
  for (int tr = 0; tr < f->tile_rows; tr++)
    for (int tc = 0; tc < f->tile_cols; tc++) {
      tile_bounds(&e, tr, tc);
      for (int r = e.tr0; r < e.tr1; r += 16)
        for (int c = e.tc0; c < e.tc1; c += 16)
          for (int z = 0; z < 256; z++) {          /* 4x4 units of the superblock in z-order */
            int ur = 0, uc = 0;
            for (int b = 0; b < 4; b++) { uc |= ((z >> (2 * b)) & 1) << b; ur |= ((z >> (2 * b + 1)) & 1) << b; }
            const int mr = r + ur, mc = c + uc;
            if (mr >= f->mi_rows || mc >= f->mi_cols) continue;
            const int n = 4 * k_bw4[mi[mr * f->mi_cols + mc].bsize];
            if ((mr * 4) % n || (mc * 4) % n) continue;      /* not the top-left unit of its block */
            reconstruct_block(&e, mc * 4, mr * 4, n);
          }
    }
  for (int lv = 0; lv < B200_LEVELS; lv++) if (!(dec_out && dec_out[lv])) free(e.dec[lv]);
}

/* test hooks: the 1-D edge operations on a caller buffer whose element [off] is the corner sample (index -1) */
void orc_ief_filter_hook(uint16_t *buf, int off, int sz, int strength) { ief_filter(buf + off + 1, sz, strength); }
void orc_ief_upsample_hook(uint16_t *buf, int off, int num_px, int bd) { ief_upsample_edge(buf + off + 1, num_px, bd); }
int orc_ief_strength_hook(int blk_wh, int filter_type, int delta) { return ief_strength(blk_wh, filter_type, delta); }
int orc_ief_use_upsample_hook(int blk_wh, int filter_type, int delta) { return ief_upsample(blk_wh, filter_type, delta); }
