/* ORACLE (test infrastructure).  2-D AV1 integer transforms, square sizes 4..32.
 *
 * Inverse: AV1 normative reconstruction (row pass, intermediate clamp, column pass, add+clip);
 * pinned against libaom 3.13.1 av1_inv_txfm2d_add_NxN_c and against dav1d through the
 * bitstream.  Forward: encoder-side; follows the libaom av1_fwd_txfm2d design (column pass
 * with input up-shift, row pass, per-size shifts); pinned against av1_fwd_txfm2d_NxN_c.
 * rav1e 0.8.1 (the reference's codec dependency, ravif/Cargo.toml:19) is not on disk, so
 * bit-parity with rav1e's forward path is "parity unpinned" (SURVEY.md section 0.3).
 *
 * Coefficient layout everywhere in this repo: row-major, coef[r*n + c], r = vertical
 * frequency, c = horizontal frequency (AV1 specification layout).
 */
#include <stdint.h>
#include <string.h>
#include "../cavif_rs_b200/csrc/av1_tables.h"
#include "txfm1d_gen.h"

enum { DCT_DCT = 0, ADST_DCT = 1, DCT_ADST = 2, ADST_ADST = 3, IDTX = 9, V_DCT = 10, H_DCT = 11 };
enum { T_DCT = 0, T_ADST = 1, T_IDT = 2 };

/* vertical (column) kernel, horizontal (row) kernel for a tx_type */
static void tx_kernels(int tx_type, int *vert, int *horz) {
  switch (tx_type) {
    case DCT_DCT: *vert = T_DCT; *horz = T_DCT; break;
    case ADST_DCT: *vert = T_ADST; *horz = T_DCT; break;
    case DCT_ADST: *vert = T_DCT; *horz = T_ADST; break;
    case ADST_ADST: *vert = T_ADST; *horz = T_ADST; break;
    case V_DCT: *vert = T_DCT; *horz = T_IDT; break;      /* DCT down the columns, identity along the rows */
    case H_DCT: *vert = T_IDT; *horz = T_DCT; break;
    default: *vert = T_IDT; *horz = T_IDT; break;
  }
}

static inline int32_t round_shift64(int64_t v, int bit) { return (int32_t)((v + ((int64_t)1 << (bit - 1))) >> bit); }
static inline int32_t rshift_round(int32_t v, int bit) { return bit <= 0 ? (int32_t)((uint32_t)v << (-bit)) : (int32_t)(((int64_t)v + ((int64_t)1 << (bit - 1))) >> bit); }

/* 4-point ADST (sinpi based), forward and inverse */
static void fadst4(const int32_t *in, int32_t *out, int bit) {
  const int32_t *sinpi = av1_sinpi_arr[bit - 10];
  int64_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
  if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  int64_t s0 = sinpi[1] * x0, s1 = sinpi[4] * x0, s2 = sinpi[2] * x1, s3 = sinpi[1] * x1;
  int64_t s4 = sinpi[3] * x2, s5 = sinpi[4] * x3, s6 = sinpi[2] * x3, s7 = x0 + x1 - x3;
  x0 = s0 + s2 + s5; x1 = sinpi[3] * s7; x2 = s1 - s3 + s6; x3 = s4;
  s0 = x0 + x3; s1 = x1; s2 = x2 - x3; s3 = x2 - x0 + x3;
  out[0] = round_shift64(s0, bit); out[1] = round_shift64(s1, bit);
  out[2] = round_shift64(s2, bit); out[3] = round_shift64(s3, bit);
}
static void iadst4(const int32_t *in, int32_t *out, int bit) {
  const int32_t *sinpi = av1_sinpi_arr[bit - 10];
  int64_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
  if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  int64_t s0 = sinpi[1] * x0, s1 = sinpi[2] * x0, s2 = sinpi[3] * x1, s3 = sinpi[4] * x2;
  int64_t s4 = sinpi[1] * x2, s5 = sinpi[2] * x3, s6 = sinpi[4] * x3, s7 = (x0 - x2) + x3;
  s0 = s0 + s3; s1 = s1 - s4; s3 = s2; s2 = sinpi[3] * s7;
  s0 = s0 + s5; s1 = s1 - s6;
  x0 = s0 + s3; x1 = s1 + s3; x2 = s2; x3 = s0 + s1; x3 = x3 - s3;
  out[0] = round_shift64(x0, bit); out[1] = round_shift64(x1, bit);
  out[2] = round_shift64(x2, bit); out[3] = round_shift64(x3, bit);
}

#define NEW_SQRT2 5793 /* 2^12 * sqrt(2) */
static void identity1d(const int32_t *in, int32_t *out, int n) {
  for (int i = 0; i < n; i++) {
    switch (n) {
      case 4: out[i] = round_shift64((int64_t)in[i] * NEW_SQRT2, 12); break;
      case 8: out[i] = in[i] * 2; break;
      case 16: out[i] = round_shift64((int64_t)in[i] * 2 * NEW_SQRT2, 12); break;
      default: out[i] = in[i] * 4; break;
    }
  }
}

static void fwd1d(int kind, int n, const int32_t *in, int32_t *out, int bit) {
  const int32_t *c = av1_cospi_arr[bit - 10];
  if (kind == T_IDT) { identity1d(in, out, n); return; }
  if (kind == T_DCT) {
    switch (n) { case 4: gen_fdct4(in, out, c, bit); break; case 8: gen_fdct8(in, out, c, bit); break;
                 case 16: gen_fdct16(in, out, c, bit); break; default: gen_fdct32(in, out, c, bit); }
    return;
  }
  switch (n) { case 4: fadst4(in, out, bit); break; case 8: gen_fadst8(in, out, c, bit); break;
               default: gen_fadst16(in, out, c, bit); }
}
static void inv1d(int kind, int n, const int32_t *in, int32_t *out) {
  const int bit = 12;
  const int32_t *c = av1_cospi_arr[bit - 10];
  if (kind == T_IDT) { identity1d(in, out, n); return; }
  if (kind == T_DCT) {
    switch (n) { case 4: gen_idct4(in, out, c, bit); break; case 8: gen_idct8(in, out, c, bit); break;
                 case 16: gen_idct16(in, out, c, bit); break; default: gen_idct32(in, out, c, bit); }
    return;
  }
  switch (n) { case 4: iadst4(in, out, bit); break; case 8: gen_iadst8(in, out, c, bit); break;
               default: gen_iadst16(in, out, c, bit); }
}

static int log2n(int n) { return n == 4 ? 2 : n == 8 ? 3 : n == 16 ? 4 : 5; }

/* libaom fwd_txfm_shift_ls / fwd_cos_bit_{col,row} for square sizes (index log2n-2) */
static const int8_t k_fwd_shift[4][3] = {{2, 0, 0}, {2, -1, 0}, {2, -2, 0}, {2, -4, 0}};
static const int8_t k_fwd_cos_col[4] = {13, 13, 13, 12};
static const int8_t k_fwd_cos_row[4] = {13, 13, 12, 12};

/* residual: n x n int16, stride in elements.  coef out: row-major n*n int32. */
void orc_fwd_txfm2d(const int16_t *resid, int stride, int32_t *coef, int n, int tx_type) {
  int vert, horz;
  tx_kernels(tx_type, &vert, &horz);
  const int s = log2n(n) - 2;
  int32_t buf[32 * 32], tin[32], tout[32];
  for (int c = 0; c < n; c++) {
    for (int r = 0; r < n; r++) tin[r] = rshift_round(resid[r * stride + c], -k_fwd_shift[s][0]);
    fwd1d(vert, n, tin, tout, k_fwd_cos_col[s]);
    for (int r = 0; r < n; r++) buf[r * n + c] = rshift_round(tout[r], -k_fwd_shift[s][1]);
  }
  for (int r = 0; r < n; r++) {
    fwd1d(horz, n, buf + r * n, tout, k_fwd_cos_row[s]);
    for (int c = 0; c < n; c++) coef[r * n + c] = rshift_round(tout[c], -k_fwd_shift[s][2]);
  }
}

static inline int32_t clamp_bits(int64_t v, int bits) {
  const int64_t mx = ((int64_t)1 << (bits - 1)) - 1, mn = -((int64_t)1 << (bits - 1));
  return (int32_t)(v < mn ? mn : v > mx ? mx : v);
}

/* AV1 spec 7.13.3 (2-D inverse transform) for square sizes; dst holds the prediction and
 * receives the reconstruction.  coef: row-major dequantised coefficients. */
void orc_inv_txfm2d_add(const int32_t *coef, uint16_t *dst, int stride, int n, int tx_type, int bd) {
  static const int8_t row_shift[4] = {0, 1, 2, 2};
  int vert, horz;
  tx_kernels(tx_type, &vert, &horz);
  const int s = log2n(n) - 2;
  int32_t buf[32 * 32], tin[32], tout[32];
  const int inter_bits = bd + 8 > 16 ? bd + 8 : 16;   /* Max(BitDepth + 6, 16) + 2?  see note */
  (void)inter_bits;
  for (int r = 0; r < n; r++) {
    for (int c = 0; c < n; c++) tin[c] = clamp_bits(coef[r * n + c], bd + 8);
    inv1d(horz, n, tin, tout);
    for (int c = 0; c < n; c++) buf[r * n + c] = rshift_round(tout[c], row_shift[s]);
  }
  const int col_clamp = (bd + 6 > 16) ? bd + 6 : 16;
  for (int c = 0; c < n; c++) {
    for (int r = 0; r < n; r++) tin[r] = clamp_bits(buf[r * n + c], col_clamp);
    inv1d(vert, n, tin, tout);
    for (int r = 0; r < n; r++) {
      const int32_t v = (int32_t)dst[r * stride + c] + rshift_round(tout[r], 4);
      const int32_t mx = (1 << bd) - 1;
      dst[r * stride + c] = (uint16_t)(v < 0 ? 0 : v > mx ? mx : v);
    }
  }
}
