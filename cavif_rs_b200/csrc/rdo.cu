// AV1 intra rate-distortion search on the device: one CTA per TILE (tiles are the unit with no
// cross dependencies: contexts reset, no prediction across tile edges), superblocks in raster
// order inside the CTA, one WARP per candidate (prediction mode / transform type trial):
// predict -> residual -> forward transform -> quantise -> static-CDF rate -> dequantise ->
// inverse transform + reconstruct -> SSE.  Transforms are integer butterflies held in
// registers, one lane per row/column, with shared-memory staging between the two passes.
//
// Replaces rav1e's encode_tile / rdo_mode_decision / encode_tx_block [EXT: crates.io
// rav1e 0.8.1], which the reference reaches only through encode_to_av1
// (/root/reference/ravif/src/av1encoder.rs:749-771).  The bit-exact CPU restatement of THIS
// file is oracle/av1_encode.c; tests/test_rdo_parity.py compares records and reconstruction.
//
// No tensor cores on purpose: the work is 4..16-point integer butterflies with normative
// rounding at every stage (BASELINE.json north_star).
#ifdef B200_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "records.h"
#include "rdo.h"

#define AV1_TABLE(type, name, dims) __constant__ const type name dims
#include "av1_tables_dev.h"
#undef AV1_TABLE

#define TXFM_FN __device__ __forceinline__
#include "txfm1d_gen.h"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int EOFF = 16;          // edge arrays: index -1 .. 2n-1 stored at EOFF-1 ..
constexpr int kEdge = EOFF + 34;  // n <= 16
constexpr unsigned FULL = 0xffffffffu;

__constant__ const uint8_t dv_sm_weights[4 + 8 + 16] = {255, 149, 85, 64, 255, 197, 146, 105, 73, 50, 37, 32,
                                                        255, 225, 196, 170, 145, 123, 102, 84, 68, 54, 43, 33, 26, 20, 17, 16};
__constant__ const uint8_t dv_intra_mode_context[13] = {0, 1, 2, 3, 4, 4, 4, 4, 3, 0, 1, 2, 0};
__constant__ const uint8_t dv_tx_types[5] = {TX_DCT_DCT, TX_ADST_DCT, TX_DCT_ADST, TX_ADST_ADST, TX_IDTX};
__constant__ const uint8_t dv_bw4[22] = {1, 1, 2, 2, 2, 4, 4, 4, 8, 8, 8, 16, 16, 16, 32, 32, 1, 4, 2, 8, 4, 16};
__constant__ const uint8_t dv_bh4[22] = {1, 2, 1, 2, 4, 2, 4, 8, 4, 8, 16, 8, 16, 32, 16, 32, 4, 1, 8, 2, 16, 4};

struct WarpScratch {
  uint16_t pred[256];
  int32_t buf[16 * 17];        // transform staging, row stride n + 1
  int16_t qc[2][256];          // two result sets (a warp may run two candidates)
  uint16_t recon[2][256];
  uint8_t lev[20 * 20];        // padded |level| map for the rate contexts
};

struct Cand {
  long long cost;
  B200Mi mi;
  uint16_t eob[3];
  int16_t qc[3][256];
  uint16_t recon[3][256];
};

struct Shared {
  WarpScratch ws[kWarps];
  Cand cur, held8, held16;
  uint16_t src[3][256];
  uint16_t above[3][kEdge], left[3][kEdge];
  uint8_t scan[336], iscan[336];
  uint32_t satd[2][13];
  // per-task results of the full trials
  long long taskJ[10];
  uint32_t taskD[4], taskR[4];
  uint16_t taskEob[10], taskEobC[4];
  int sel[8];                  // small broadcast area
  int have_above, have_left;
};

struct TileCtx {
  const B200Frame *f;
  const B200RateTables *rt;
  const uint16_t *src[3];
  uint16_t *rec[3];
  B200Mi *mi;
  uint16_t *eob[3];
  int16_t *coef[3];
  uint8_t *decoded;
  int stride;
  int tr0, tr1, tc0, tc1;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ int txs_of(int n) { return n == 4 ? 0 : n == 8 ? 1 : n == 16 ? 2 : 3; }
__device__ __forceinline__ int scan_base(int n) { return n == 4 ? 0 : n == 8 ? 16 : 80; }
__device__ __forceinline__ int ilog2_dev(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
  for (int m = 16; m > 0; m >>= 1) v = max(v, __shfl_xor_sync(FULL, v, m));
  return v;
}

// ---- edges (AV1 spec 7.11.2, intra edge filter disabled) : executed by one warp ------------
__device__ void build_edges(const TileCtx &t, Shared &sh, int plane, int x, int y, int n) {
  const B200Frame *f = t.f;
  const uint16_t *r = t.rec[plane];
  const int st = t.stride, lane = lane_id();
  const int mi_r = y >> 2, mi_c = x >> 2, n4 = n >> 2;
  const int ha = mi_r > t.tr0, hl = mi_c > t.tc0;
  const int max_x = f->mi_cols * 4 - 1, max_y = f->mi_rows * 4 - 1;
  const int base = 1 << (f->bit_depth - 1);
  int have_ar = 0, have_bl = 0;
  if (ha && mi_c + n4 < t.tc1) have_ar = t.decoded[(mi_r - 1) * f->mi_cols + mi_c + n4];
  if (hl && mi_r + n4 < t.tr1) have_bl = t.decoded[(mi_r + n4) * f->mi_cols + mi_c - 1];
  uint16_t *A = sh.above[plane] + EOFF, *L = sh.left[plane] + EOFF;
  for (int i = lane; i < 2 * n; i += 32) {
    uint16_t a, l;
    if (!ha && hl) a = r[y * st + x - 1];
    else if (!ha && !hl) a = (uint16_t)(base - 1);
    else {
      int xx = (i < n || have_ar) ? x + i : x + n - 1;
      if (xx > max_x) xx = max_x;
      a = r[(y - 1) * st + xx];
    }
    if (!hl && ha) l = r[(y - 1) * st + x];
    else if (!hl && !ha) l = (uint16_t)(base + 1);
    else {
      int yy = (i < n || have_bl) ? y + i : y + n - 1;
      if (yy > max_y) yy = max_y;
      l = r[yy * st + x - 1];
    }
    A[i] = a; L[i] = l;
  }
  if (lane == 0) {
    uint16_t tl;
    if (ha && hl) tl = r[(y - 1) * st + x - 1];
    else if (ha) tl = r[(y - 1) * st + x];
    else if (hl) tl = r[y * st + x - 1];
    else tl = (uint16_t)base;
    A[-1] = tl; L[-1] = tl;
    if (plane == 0) { sh.have_above = ha; sh.have_left = hl; }
  }
}

// ---- intra predictors (AV1 spec 7.11.2.x), warp-cooperative over the n*n pixels -------------
__device__ void predict_intra(int mode, int n, const uint16_t *above, const uint16_t *left, int ha, int hl, int bd,
                              uint16_t *pred) {
  const uint16_t *A = above + EOFF, *L = left + EOFF;
  const int lane = lane_id(), n2 = n * n, lg = n == 4 ? 2 : n == 8 ? 3 : 4;
  if (mode == DC_PRED) {
    int v;
    if (ha && hl) { int s = 0; for (int i = 0; i < n; i++) s += A[i] + L[i]; v = (s + n) / (2 * n); }
    else if (hl) { int s = 0; for (int i = 0; i < n; i++) s += L[i]; v = (s + (n >> 1)) / n; }
    else if (ha) { int s = 0; for (int i = 0; i < n; i++) s += A[i]; v = (s + (n >> 1)) / n; }
    else v = 1 << (bd - 1);
    for (int p = lane; p < n2; p += 32) pred[p] = (uint16_t)v;
    return;
  }
  if (mode == PAETH_PRED) {
    const int tl = A[-1];
    for (int p = lane; p < n2; p += 32) {
      const int i = p >> lg, j = p & (n - 1);
      const int b = A[j] + L[i] - tl;
      const int pl = abs(b - L[i]), pt = abs(b - A[j]), ptl = abs(b - tl);
      pred[p] = (pl <= pt && pl <= ptl) ? L[i] : (pt <= ptl ? A[j] : (uint16_t)tl);
    }
    return;
  }
  if (mode >= SMOOTH_PRED && mode <= SMOOTH_H_PRED) {
    const uint8_t *w = dv_sm_weights + (n == 4 ? 0 : n == 8 ? 4 : 12);
    for (int p = lane; p < n2; p += 32) {
      const int i = p >> lg, j = p & (n - 1);
      int v;
      if (mode == SMOOTH_PRED) v = (w[i] * A[j] + (256 - w[i]) * L[n - 1] + w[j] * L[i] + (256 - w[j]) * A[n - 1] + 256) >> 9;
      else if (mode == SMOOTH_V_PRED) v = (w[i] * A[j] + (256 - w[i]) * L[n - 1] + 128) >> 8;
      else v = (w[j] * L[i] + (256 - w[j]) * A[n - 1] + 128) >> 8;
      pred[p] = (uint16_t)v;
    }
    return;
  }
  const int p_angle = dv_mode_to_angle[mode];
  if (p_angle == 90) { for (int p = lane; p < n2; p += 32) pred[p] = A[p & (n - 1)]; return; }
  if (p_angle == 180) { for (int p = lane; p < n2; p += 32) pred[p] = L[p >> lg]; return; }
  if (p_angle < 90) {
    const int dx = dv_dr_intra_derivative[p_angle], max_base = 2 * n - 1;
    for (int p = lane; p < n2; p += 32) {
      const int i = p >> lg, j = p & (n - 1);
      const int idx = (i + 1) * dx, base = (idx >> 6) + j, shift = (idx >> 1) & 31;
      pred[p] = base < max_base ? (uint16_t)((A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5) : A[max_base];
    }
  } else if (p_angle < 180) {
    const int dx = dv_dr_intra_derivative[180 - p_angle], dy = dv_dr_intra_derivative[p_angle - 90];
    for (int p = lane; p < n2; p += 32) {
      const int i = p >> lg, j = p & (n - 1);
      int idx = (j << 6) - (i + 1) * dx, base = idx >> 6;
      if (base >= -1) {
        const int shift = (idx >> 1) & 31;
        pred[p] = (uint16_t)((A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5);
      } else {
        idx = (i << 6) - (j + 1) * dy; base = idx >> 6;
        const int shift = (idx >> 1) & 31;
        pred[p] = (uint16_t)((L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5);
      }
    }
  } else {
    const int dy = dv_dr_intra_derivative[270 - p_angle], max_base = 2 * n - 1;
    for (int p = lane; p < n2; p += 32) {
      const int i = p >> lg, j = p & (n - 1);
      const int idx = (j + 1) * dy, base = (idx >> 6) + i, shift = (idx >> 1) & 31;
      pred[p] = base < max_base ? (uint16_t)((L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5) : L[max_base];
    }
  }
}

// sum of |4x4 Hadamard| over the 4x4 sub-blocks; one lane per sub-block, warp reduced
__device__ uint32_t satd4_warp(const uint16_t *src, const uint16_t *pred, int n) {
  const int nb = n >> 2, lane = lane_id();
  uint32_t total = 0;
  for (int s = lane; s < nb * nb; s += 32) {
    const int by = (s / nb) * 4, bx = (s % nb) * 4;
    int d[16], t[16];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) d[i * 4 + j] = (int)src[(by + i) * n + bx + j] - (int)pred[(by + i) * n + bx + j];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int a = d[i * 4] + d[i * 4 + 1], b = d[i * 4] - d[i * 4 + 1], c = d[i * 4 + 2] + d[i * 4 + 3], e = d[i * 4 + 2] - d[i * 4 + 3];
      t[i * 4] = a + c; t[i * 4 + 1] = b + e; t[i * 4 + 2] = a - c; t[i * 4 + 3] = b - e;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int a = t[j] + t[4 + j], b = t[j] - t[4 + j], c = t[8 + j] + t[12 + j], e = t[8 + j] - t[12 + j];
      total += (uint32_t)(abs(a + c) + abs(b + e) + abs(a - c) + abs(b - e));
    }
  }
  return warp_sum(total);
}

// ---- 1-D transforms on register arrays -----------------------------------------------------
enum { K_DCT = 0, K_ADST = 1, K_IDT = 2 };
__device__ __forceinline__ void tx_kernels(int tx_type, int &vert, int &horz) {
  vert = (tx_type == TX_ADST_DCT || tx_type == TX_ADST_ADST) ? K_ADST : (tx_type == TX_IDTX ? K_IDT : K_DCT);
  horz = (tx_type == TX_DCT_ADST || tx_type == TX_ADST_ADST) ? K_ADST : (tx_type == TX_IDTX ? K_IDT : K_DCT);
}
__device__ __forceinline__ int32_t round_shift64(long long v, int bit) { return (int32_t)((v + (1ll << (bit - 1))) >> bit); }

__device__ __forceinline__ void fadst4_dev(const int32_t *in, int32_t *out, int bit) {
  const int32_t *sinpi = dv_sinpi_arr[bit - 10];
  long long x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
  if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  long long s0 = sinpi[1] * x0, s1 = sinpi[4] * x0, s2 = sinpi[2] * x1, s3 = sinpi[1] * x1;
  long long s4 = sinpi[3] * x2, s5 = sinpi[4] * x3, s6 = sinpi[2] * x3, s7 = x0 + x1 - x3;
  x0 = s0 + s2 + s5; x1 = sinpi[3] * s7; x2 = s1 - s3 + s6; x3 = s4;
  s0 = x0 + x3; s1 = x1; s2 = x2 - x3; s3 = x2 - x0 + x3;
  out[0] = round_shift64(s0, bit); out[1] = round_shift64(s1, bit);
  out[2] = round_shift64(s2, bit); out[3] = round_shift64(s3, bit);
}
__device__ __forceinline__ void iadst4_dev(const int32_t *in, int32_t *out, int bit) {
  const int32_t *sinpi = dv_sinpi_arr[bit - 10];
  long long x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
  if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  long long s0 = sinpi[1] * x0, s1 = sinpi[2] * x0, s2 = sinpi[3] * x1, s3 = sinpi[4] * x2;
  long long s4 = sinpi[1] * x2, s5 = sinpi[2] * x3, s6 = sinpi[4] * x3, s7 = (x0 - x2) + x3;
  s0 = s0 + s3; s1 = s1 - s4; s3 = s2; s2 = sinpi[3] * s7;
  s0 = s0 + s5; s1 = s1 - s6;
  x0 = s0 + s3; x1 = s1 + s3; x2 = s2; x3 = s0 + s1; x3 = x3 - s3;
  out[0] = round_shift64(x0, bit); out[1] = round_shift64(x1, bit);
  out[2] = round_shift64(x2, bit); out[3] = round_shift64(x3, bit);
}
template <int N>
__device__ __forceinline__ void identity_dev(const int32_t *in, int32_t *out) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (N == 4) out[i] = round_shift64((long long)in[i] * 5793, 12);
    else if (N == 8) out[i] = in[i] * 2;
    else out[i] = round_shift64((long long)in[i] * 2 * 5793, 12);
  }
}
template <int N>
__device__ __forceinline__ void fwd1d(int kind, const int32_t *in, int32_t *out, int bit) {
  const int32_t *c = dv_cospi_arr[bit - 10];
  if (kind == K_IDT) { identity_dev<N>(in, out); return; }
  if (kind == K_DCT) {
    if (N == 4) gen_fdct4(in, out, c, bit); else if (N == 8) gen_fdct8(in, out, c, bit); else gen_fdct16(in, out, c, bit);
  } else {
    if (N == 4) fadst4_dev(in, out, bit); else if (N == 8) gen_fadst8(in, out, c, bit); else gen_fadst16(in, out, c, bit);
  }
}
template <int N>
__device__ __forceinline__ void inv1d(int kind, const int32_t *in, int32_t *out) {
  const int bit = 12;
  const int32_t *c = dv_cospi_arr[bit - 10];
  if (kind == K_IDT) { identity_dev<N>(in, out); return; }
  if (kind == K_DCT) {
    if (N == 4) gen_idct4(in, out, c, bit); else if (N == 8) gen_idct8(in, out, c, bit); else gen_idct16(in, out, c, bit);
  } else {
    if (N == 4) iadst4_dev(in, out, bit); else if (N == 8) gen_iadst8(in, out, c, bit); else gen_iadst16(in, out, c, bit);
  }
}
__device__ __forceinline__ int32_t rshift_round(int32_t v, int bit) {   // bit > 0: rounding right shift; <= 0: left shift
  return bit <= 0 ? (int32_t)((uint32_t)v << (-bit)) : (int32_t)(((long long)v + (1ll << (bit - 1))) >> bit);
}
__device__ __forceinline__ int32_t clamp_bits(int32_t v, int bits) {
  const int32_t mx = (1 << (bits - 1)) - 1, mn = -(1 << (bits - 1));
  return v < mn ? mn : (v > mx ? mx : v);
}

// forward 2-D transform of (src - pred) into buf (row stride N+1), libaom av1_fwd_txfm2d design
template <int N>
__device__ void fwd_txfm2d_warp(const uint16_t *src, const uint16_t *pred, int32_t *buf, int tx_type) {
  constexpr int S = N + 1;
  constexpr int shift1 = N == 4 ? 0 : (N == 8 ? 1 : 2);   // -fwd_shift[1]
  constexpr int cos_col = 13, cos_row = N == 16 ? 12 : 13;
  const int lane = lane_id();
  int vert, horz;
  tx_kernels(tx_type, vert, horz);
  for (int p = lane; p < N * N; p += 32) {
    const int r = p / N, c = p % N;
    buf[r * S + c] = ((int)src[p] - (int)pred[p]) * 4;   // fwd_shift[0] = 2
  }
  __syncwarp();
  if (lane < N) {
    int32_t x[N], o[N];
#pragma unroll
    for (int r = 0; r < N; r++) x[r] = buf[r * S + lane];
    fwd1d<N>(vert, x, o, cos_col);
#pragma unroll
    for (int r = 0; r < N; r++) buf[r * S + lane] = rshift_round(o[r], shift1);
  }
  __syncwarp();
  if (lane < N) {
    int32_t x[N], o[N];
#pragma unroll
    for (int c = 0; c < N; c++) x[c] = buf[lane * S + c];
    fwd1d<N>(horz, x, o, cos_row);
#pragma unroll
    for (int c = 0; c < N; c++) buf[lane * S + c] = o[c];
  }
  __syncwarp();
}

// inverse 2-D transform (AV1 spec 7.13.3) of buf (dequantised, stride N+1) added to pred -> recon
template <int N>
__device__ void inv_txfm2d_add_warp(int32_t *buf, const uint16_t *pred, uint16_t *recon, int tx_type, int bd) {
  constexpr int S = N + 1;
  constexpr int row_shift = N == 4 ? 0 : (N == 8 ? 1 : 2);
  const int lane = lane_id();
  int vert, horz;
  tx_kernels(tx_type, vert, horz);
  if (lane < N) {
    int32_t x[N], o[N];
#pragma unroll
    for (int c = 0; c < N; c++) x[c] = clamp_bits(buf[lane * S + c], bd + 8);
    inv1d<N>(horz, x, o);
#pragma unroll
    for (int c = 0; c < N; c++) buf[lane * S + c] = rshift_round(o[c], row_shift);
  }
  __syncwarp();
  if (lane < N) {
    const int col_clamp = bd + 6 > 16 ? bd + 6 : 16;
    const int mx = (1 << bd) - 1;
    int32_t x[N], o[N];
#pragma unroll
    for (int r = 0; r < N; r++) x[r] = clamp_bits(buf[r * S + lane], col_clamp);
    inv1d<N>(vert, x, o);
#pragma unroll
    for (int r = 0; r < N; r++) {
      const int v = (int)pred[r * N + lane] + rshift_round(o[r], 4);
      recon[r * N + lane] = (uint16_t)(v < 0 ? 0 : (v > mx ? mx : v));
    }
  }
  __syncwarp();
}

// Full trial of one plane by one warp.  src/pred: n*n in shared memory.  Results: qc (scan order),
// recon, *eob, returns SSE, *rate.  Mirrors oracle/av1_encode.c trial_plane().
template <int N>
__device__ uint32_t trial_plane_warp(const TileCtx &t, Shared &sh, WarpScratch &w, int set, int plane, const uint16_t *src,
                                     int tx_type, int y_mode, int *eob_out, uint32_t *rate_out) {
  constexpr int S = N + 1, N2 = N * N, LS = N + 4;
  const B200Frame *f = t.f;
  const B200RateTables *rt = t.rt;
  const int lane = lane_id();
  const uint8_t *iscan = sh.iscan + scan_base(N);
  const uint16_t *pred = w.pred;
  int16_t *qc = w.qc[set];
  fwd_txfm2d_warp<N>(src, pred, w.buf, tx_type);
  // ---- quantise / dequantise (oracle quantize())
  for (int p = lane; p < LS * LS; p += 32) w.lev[p] = 0;
  __syncwarp();
  int eob = 0;
  const int dc_q = f->dc_q, ac_q = f->ac_q;
  for (int pos = lane; pos < N2; pos += 32) {
    const int r = pos / N, c = pos % N;
    const int q = pos == 0 ? dc_q : ac_q;
    const int off = pos == 0 ? (q * 109) >> 8 : (q * 98) >> 8;
    const int32_t cf = w.buf[r * S + c];
    const uint32_t a = (uint32_t)(cf < 0 ? -cf : cf);
    uint32_t lv = (a + (uint32_t)off) / (uint32_t)q;
    if (lv > 32767) lv = 32767;
    const int32_t dqa = (int32_t)(((unsigned long long)lv * (uint32_t)q) & 0xFFFFFF);
    const int si = iscan[pos];
    qc[si] = (int16_t)(cf < 0 ? -(int)lv : (int)lv);
    w.buf[r * S + c] = cf < 0 ? -dqa : dqa;
    w.lev[r * LS + c] = (uint8_t)(lv > 127 ? 127 : lv);
    if (lv) eob = max(eob, si + 1);
  }
  eob = warp_max(eob);
  __syncwarp();
  // ---- rate (oracle coef_rate())
  const int txs = txs_of(N), ptype = plane > 0;
  const int skip_ctx = plane == 0 ? 0 : 7;
  uint32_t rate = 0;
  if (eob == 0) {
    rate = rt->txb_skip[txs][skip_ctx][1];
  } else {
    if (lane == 0) {
      rate = rt->txb_skip[txs][skip_ctx][0];
      if (plane == 0) rate += rt->intra_tx[(f->reduced_tx_set || N == 16) ? 1 : 0][txs][y_mode][tx_type];
      int tg = 0;
      const int group_start[12] = {0, 1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 513};
      while (tg < 11 && group_start[tg + 1] <= eob) tg++;
      rate += rt->eob_pt[txs][ptype][tg - 1];
      const int nb = tg < 3 ? 0 : tg - 2;
      if (nb > 0) {
        const int extra = eob - group_start[tg];
        rate += rt->eob_extra[txs][ptype][tg - 3][(extra >> (nb - 1)) & 1];
        rate += 256u * (uint32_t)(nb - 1);
      }
    }
    for (int pos = lane; pos < N2; pos += 32) {
      const int si = iscan[pos];
      if (si >= eob) continue;
      const int rr = pos / N, cc = pos % N;
      const int level = abs((int)qc[si]);
      const uint8_t *L = w.lev + rr * LS + cc;
      if (si == eob - 1) {
        const int c4 = si == 0 ? 0 : (si <= N2 / 8 ? 1 : (si <= N2 / 4 ? 2 : 3));
        rate += rt->coeff_base_eob[txs][ptype][c4][min(level, 3) - 1];
      } else {
        const int mag = min((int)L[1], 3) + min((int)L[LS], 3) + min((int)L[LS + 1], 3) + min((int)L[2], 3) + min((int)L[2 * LS], 3);
        int ctx = min((mag + 1) >> 1, 4);
        if (rr == 0 && cc == 0) ctx = 0;
        else ctx += (rr + cc < 2) ? 1 : ((rr + cc < 4) ? 6 : 21);
        rate += rt->coeff_base[txs][ptype][ctx][min(level, 3)];
      }
      if (level > 2) {
        int mag = min((int)L[1], 15) + min((int)L[LS], 15) + min((int)L[LS + 1], 15);
        mag = min((mag + 1) >> 1, 6);
        const int bctx = (rr == 0 && cc == 0) ? mag : ((rr < 2 && cc < 2) ? mag + 7 : mag + 14);
        rate += rt->coeff_br[txs][ptype][bctx][min(level - 3, 12)];
        if (level > 14) {
          int xg = level - 15 + 1, len = 0;
          while (xg) { xg >>= 1; len++; }
          rate += 256u * (uint32_t)(2 * len - 1);
        }
      }
      if (level) rate += (si == 0) ? rt->dc_sign[ptype][0][qc[si] < 0] : 256u;
    }
    rate = warp_sum(rate);
  }
  // ---- reconstruct
  if (eob) {
    inv_txfm2d_add_warp<N>(w.buf, pred, w.recon[set], tx_type, f->bit_depth);
  } else {
    for (int p = lane; p < N2; p += 32) w.recon[set][p] = pred[p];
    __syncwarp();
  }
  // ---- distortion
  uint32_t sse = 0;
  for (int p = lane; p < N2; p += 32) {
    const int d = (int)src[p] - (int)w.recon[set][p];
    sse += (uint32_t)(d * d);
  }
  sse = warp_sum(sse);
  *eob_out = eob;
  *rate_out = rate;
  return sse;
}

__device__ __forceinline__ long long rd_cost(const B200Frame *f, unsigned long long sse, uint32_t rate) {
  return (long long)(sse << 8) + (((long long)f->lambda_q8 * (long long)rate) >> 8);
}

__device__ __forceinline__ uint32_t trial_plane_dispatch(int n, const TileCtx &t, Shared &sh, WarpScratch &w, int set, int plane,
                                                         const uint16_t *src, int tx_type, int y_mode, int *eob,
                                                         uint32_t *rate) {
  if (n == 4) return trial_plane_warp<4>(t, sh, w, set, plane, src, tx_type, y_mode, eob, rate);
  if (n == 8) return trial_plane_warp<8>(t, sh, w, set, plane, src, tx_type, y_mode, eob, rate);
  return trial_plane_warp<16>(t, sh, w, set, plane, src, tx_type, y_mode, eob, rate);
}

__device__ __forceinline__ int chroma_tx_type(int uv_mode) { return dv_intra_mode_to_tx_type[uv_mode]; }

__device__ __forceinline__ uint32_t mode_rate(const B200RateTables *rt, const uint16_t *cost_row, int m, int n) {
  uint32_t r = cost_row[m];
  if (n >= 8 && m >= V_PRED && m <= D67_PRED) r += rt->angle_delta[m - V_PRED][3];
  return r;
}

// Evaluate one block (all planes) at (x, y), size n (4|8|16): result in sh.cur.  All threads of the CTA.
__device__ void trial_block(const TileCtx &t, Shared &sh, int x, int y, int n) {
  const B200Frame *f = t.f;
  const B200RateTables *rt = t.rt;
  const int tid = threadIdx.x, wid = warp_id(), lane = lane_id();
  const int n2 = n * n, planes = f->planes;
  const int mi_r = y >> 2, mi_c = x >> 2;
  const int avail_u = mi_r > t.tr0, avail_l = mi_c > t.tc0;
  // ---- P0: stage source block, build edges
  for (int i = tid; i < planes * n2; i += kThreads) {
    const int p = i / n2, q = i % n2;
    sh.src[p][q] = t.src[p][(y + q / n) * t.stride + x + (q % n)];
  }
  if (wid < planes) build_edges(t, sh, wid, x, y, n);
  __syncthreads();
  const int ha = sh.have_above, hl = sh.have_left, bd = f->bit_depth;
  // ---- P1: SATD of every allowed mode: luma tasks 0..12, chroma (U+V) tasks 13..25
  {
    WarpScratch &w = sh.ws[wid];
    const int ntask = planes > 1 ? 26 : 13;
    for (int task = wid; task < ntask; task += kWarps) {
      const int g = task / 13, m = task % 13;
      const int mask = g == 0 ? f->y_mode_mask : f->uv_mode_mask;
      uint32_t s = 0xffffffffu;
      if ((mask >> m) & 1) {
        if (g == 0) {
          predict_intra(m, n, sh.above[0], sh.left[0], ha, hl, bd, w.pred);
          __syncwarp();
          s = satd4_warp(sh.src[0], w.pred, n);
        } else {
          predict_intra(m, n, sh.above[1], sh.left[1], ha, hl, bd, w.pred);
          __syncwarp();
          s = satd4_warp(sh.src[1], w.pred, n);
          __syncwarp();
          predict_intra(m, n, sh.above[2], sh.left[2], ha, hl, bd, w.pred);
          __syncwarp();
          s += satd4_warp(sh.src[2], w.pred, n);
        }
        __syncwarp();
      }
      if (lane == 0) sh.satd[g][m] = s;
    }
  }
  __syncthreads();
  // ---- P2: rank luma modes (thread 0)
  const int am = avail_u ? t.mi[(mi_r - 1) * f->mi_cols + mi_c].y_mode : DC_PRED;
  const int lm = avail_l ? t.mi[mi_r * f->mi_cols + mi_c - 1].y_mode : DC_PRED;
  const uint16_t *ymode_cost = rt->y_mode[dv_intra_mode_context[am]][dv_intra_mode_context[lm]];
  if (tid == 0) {
    long long bc0 = 0x7fffffffffffffffll, bc1 = 0x7fffffffffffffffll;
    int bm0 = -1, bm1 = -1;
    for (int m = 0; m < 13; m++) {
      if (!((f->y_mode_mask >> m) & 1)) continue;
      const long long c = (long long)sh.satd[0][m] * 64 + (((long long)f->lambda_satd_q8 * mode_rate(rt, ymode_cost, m, n)) >> 8);
      if (c < bc0) { bc1 = bc0; bm1 = bm0; bc0 = c; bm0 = m; }
      else if (c < bc1) { bc1 = c; bm1 = m; }
    }
    sh.sel[0] = bm0; sh.sel[1] = bm1;
  }
  __syncthreads();
  // ---- P3: luma full trials: candidate c = k * 5 + t  (mode rank k, transform type index t)
  int first_t = -1;
  for (int k = 0; k < 5; k++) if ((f->tx_type_mask >> k) & 1) { first_t = k; break; }
  {
    WarpScratch &w = sh.ws[wid];
    int have = 0;              // number of candidates this warp already holds (kept in set 0)
    long long myJ = 0;
    for (int c = wid; c < 10; c += kWarps) {
      const int k = c / 5, ti = c % 5;
      const int m = sh.sel[k];
      const bool ok = m >= 0 && ((f->tx_type_mask >> ti) & 1) && (f->rdo_tx_decision || ti == first_t);
      if (!ok) { if (lane == 0) sh.taskJ[c] = 0x7fffffffffffffffll; continue; }
      predict_intra(m, n, sh.above[0], sh.left[0], ha, hl, bd, w.pred);
      __syncwarp();
      int eob; uint32_t rate;
      const int set = have ? 1 : 0;
      const uint32_t d = trial_plane_dispatch(n, t, sh, w, set, 0, sh.src[0], dv_tx_types[ti], m, &eob, &rate);
      const long long J = rd_cost(f, d, rate + mode_rate(rt, ymode_cost, m, n));
      if (lane == 0) { sh.taskJ[c] = J; sh.taskEob[c] = (uint16_t)eob; }
      if (have && J < myJ) {   // second candidate strictly better: move it to set 0
        for (int p = lane; p < n2; p += 32) { w.qc[0][p] = w.qc[1][p]; w.recon[0][p] = w.recon[1][p]; }
        myJ = J;
      } else if (!have) { myJ = J; have = 1; }
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- P4: pick the luma winner
  if (tid == 0) {
    long long bj = 0x7fffffffffffffffll;
    int bcand = 0;
    for (int c = 0; c < 10; c++) if (sh.taskJ[c] < bj) { bj = sh.taskJ[c]; bcand = c; }
    sh.sel[2] = bcand;
  }
  __syncthreads();
  const int bcand = sh.sel[2];
  const int best_mode = sh.sel[bcand / 5], best_tt = dv_tx_types[bcand % 5];
  const long long lumaJ = sh.taskJ[bcand];
  {
    // the winner's data sit in set 0 of warp (bcand % kWarps)
    const WarpScratch &w = sh.ws[bcand % kWarps];
    for (int p = tid; p < n2; p += kThreads) { sh.cur.qc[0][p] = w.qc[0][p]; sh.cur.recon[0][p] = w.recon[0][p]; }
    if (tid == 0) sh.cur.eob[0] = sh.taskEob[bcand];
  }
  long long total = lumaJ;
  int uv_best = DC_PRED;
  if (planes > 1) {
    // ---- P5: rank uv modes
    const uint16_t *uvcost = rt->uv_mode[1][best_mode];   // cfl_allowed = 1 for n <= 32
    if (tid == 0) {
      long long bc0 = 0x7fffffffffffffffll, bc1 = 0x7fffffffffffffffll;
      int bm0 = -1, bm1 = -1;
      for (int m = 0; m < 13; m++) {
        if (!((f->uv_mode_mask >> m) & 1)) continue;
        const long long c = (long long)sh.satd[1][m] * 64 + (((long long)f->lambda_satd_q8 * mode_rate(rt, uvcost, m, n)) >> 8);
        if (c < bc0) { bc1 = bc0; bm1 = bm0; bc0 = c; bm0 = m; }
        else if (c < bc1) { bc1 = c; bm1 = m; }
      }
      sh.sel[3] = bm0; sh.sel[4] = bm1;
    }
    __syncthreads();   // also orders the copy of the luma winner before warps reuse their scratch
    // ---- P6: chroma full trials: task = k * 2 + (plane - 1), warps 0..3
    if (wid < 4) {
      WarpScratch &w = sh.ws[wid];
      const int k = wid >> 1, plane = 1 + (wid & 1);
      const int m = sh.sel[3 + k];
      if (m >= 0) {
        predict_intra(m, n, sh.above[plane], sh.left[plane], ha, hl, bd, w.pred);
        __syncwarp();
        int eob; uint32_t rate;
        const uint32_t d = trial_plane_dispatch(n, t, sh, w, 0, plane, sh.src[plane], chroma_tx_type(m), best_mode, &eob, &rate);
        if (lane == 0) { sh.taskD[wid] = d; sh.taskR[wid] = rate; sh.taskEobC[wid] = (uint16_t)eob; }
      }
    }
    __syncthreads();
    // ---- P7: pick the chroma winner
    long long bestC = 0x7fffffffffffffffll;
    int bk = 0;
    for (int k = 0; k < 2; k++) {
      const int m = sh.sel[3 + k];
      if (m < 0) continue;
      const long long J = rd_cost(f, (unsigned long long)sh.taskD[2 * k] + sh.taskD[2 * k + 1],
                                  sh.taskR[2 * k] + sh.taskR[2 * k + 1] + mode_rate(rt, uvcost, m, n));
      if (J < bestC) { bestC = J; bk = k; }
    }
    uv_best = sh.sel[3 + bk];
    total += bestC;
    for (int p = tid; p < n2; p += kThreads) {
      sh.cur.qc[1][p] = sh.ws[2 * bk].qc[0][p]; sh.cur.recon[1][p] = sh.ws[2 * bk].recon[0][p];
      sh.cur.qc[2][p] = sh.ws[2 * bk + 1].qc[0][p]; sh.cur.recon[2][p] = sh.ws[2 * bk + 1].recon[0][p];
    }
    if (tid == 0) { sh.cur.eob[1] = sh.taskEobC[2 * bk]; sh.cur.eob[2] = sh.taskEobC[2 * bk + 1]; }
  }
  // ---- skip flag + totals
  if (tid == 0) {
    B200Mi m;
    m.bsize = (uint8_t)(n == 4 ? BLOCK_4X4 : n == 8 ? BLOCK_8X8 : BLOCK_16X16);
    m.y_mode = (uint8_t)best_mode;
    m.uv_mode = (uint8_t)uv_best;
    m.angle = 3 | (3 << 4);
    m.tx_size = (uint8_t)txs_of(n);
    m.tx_type = (uint8_t)best_tt;
    m.flags = 0;
    m.cfl = 0;
    const int all_zero = sh.cur.eob[0] == 0 && (planes == 1 || (sh.cur.eob[1] == 0 && sh.cur.eob[2] == 0));
    const int a = avail_u ? (t.mi[(mi_r - 1) * f->mi_cols + mi_c].flags & 1) : 0;
    const int l = avail_l ? (t.mi[mi_r * f->mi_cols + mi_c - 1].flags & 1) : 0;
    total += ((long long)f->lambda_q8 * rt->skip[a + l][all_zero]) >> 8;
    if (all_zero) m.flags |= 1;
    sh.cur.mi = m;
    sh.cur.cost = total;
  }
  __syncthreads();
}

__device__ void copy_cand(Cand &dst, const Cand &src, int n, int planes) {
  const int tid = threadIdx.x, n2 = n * n;
  for (int i = tid; i < planes * n2; i += kThreads) {
    const int p = i / n2, q = i % n2;
    dst.qc[p][q] = src.qc[p][q];
    dst.recon[p][q] = src.recon[p][q];
  }
  if (tid == 0) { dst.cost = src.cost; dst.mi = src.mi; dst.eob[0] = src.eob[0]; dst.eob[1] = src.eob[1]; dst.eob[2] = src.eob[2]; }
  __syncthreads();
}

__device__ void commit_block(const TileCtx &t, const Cand &cd, int x, int y, int n) {
  const B200Frame *f = t.f;
  const int tid = threadIdx.x, n2 = n * n, planes = f->planes;
  const int mi_r = y >> 2, mi_c = x >> 2, n4 = n >> 2;
  for (int i = tid; i < planes * n2; i += kThreads) {
    const int p = i / n2, q = i % n2;
    t.rec[p][(y + q / n) * t.stride + x + (q % n)] = cd.recon[p][q];
    t.coef[p][b200_coef_offset(f, mi_r, mi_c) + q] = cd.qc[p][q];
  }
  if (tid < planes) t.eob[tid][mi_r * f->mi_cols + mi_c] = cd.eob[tid];
  for (int i = tid; i < n4 * n4; i += kThreads) {
    const int idx = (mi_r + i / n4) * f->mi_cols + mi_c + (i % n4);
    t.mi[idx] = cd.mi;
    t.decoded[idx] = 1;
  }
  __syncthreads();
}

__device__ long long partition_cost(const TileCtx &t, int mi_r, int mi_c, int n, int split) {
  const B200Frame *f = t.f;
  const int bsl = ilog2_dev(n) - 2;
  const int hbs = n >> 3;
  const int has_rows = (mi_r + hbs) < f->mi_rows, has_cols = (mi_c + hbs) < f->mi_cols;
  if (!has_rows || !has_cols) return 0;
  const int au = mi_r > t.tr0, al = mi_c > t.tc0;
  const int above = au && ilog2_dev(dv_bw4[t.mi[(mi_r - 1) * f->mi_cols + mi_c].bsize]) < bsl;
  const int left = al && ilog2_dev(dv_bh4[t.mi[mi_r * f->mi_cols + mi_c - 1].bsize]) < bsl;
  const int ctx = (bsl - 1) * 4 + left * 2 + above;
  return ((long long)f->lambda_q8 * t.rt->partition[ctx][split ? 3 : 0]) >> 8;
}

// 8x8 node: NONE vs four 4x4 (oracle encode_tree at n = 8).  Returns the cost; frame state holds the winner.
__device__ long long encode_tree8(const TileCtx &t, Shared &sh, int x, int y) {
  const B200Frame *f = t.f;
  const int mi_r = y >> 2, mi_c = x >> 2;
  const bool can_split = f->partition_min < 8;
  trial_block(t, sh, x, y, 8);
  const long long j_none = sh.cur.cost + partition_cost(t, mi_r, mi_c, 8, 0);
  if (!can_split) { commit_block(t, sh.cur, x, y, 8); return j_none; }
  copy_cand(sh.held8, sh.cur, 8, f->planes);
  long long j_split = partition_cost(t, mi_r, mi_c, 8, 1);
  for (int k = 0; k < 4; k++) {
    const int xx = x + (k & 1) * 4, yy = y + (k >> 1) * 4;
    trial_block(t, sh, xx, yy, 4);
    j_split += sh.cur.cost;
    commit_block(t, sh.cur, xx, yy, 4);
  }
  if (j_none <= j_split) { commit_block(t, sh.held8, x, y, 8); return j_none; }
  return j_split;
}

// 16x16 node
__device__ long long encode_tree16(const TileCtx &t, Shared &sh, int x, int y) {
  const B200Frame *f = t.f;
  const int W = f->mi_cols * 4, H = f->mi_rows * 4;
  if (x >= W || y >= H) return 0;
  const int mi_r = y >> 2, mi_c = x >> 2;
  const bool inside = (x + 16 <= W) && (y + 16 <= H);
  const bool can_none = inside && f->partition_max >= 16;
  const bool can_split = f->partition_min < 16 || !inside;
  long long j_none = 0x7fffffffffffffffll;
  if (can_none) {
    trial_block(t, sh, x, y, 16);
    j_none = sh.cur.cost + partition_cost(t, mi_r, mi_c, 16, 0);
    if (!can_split) { commit_block(t, sh.cur, x, y, 16); return j_none; }
    copy_cand(sh.held16, sh.cur, 16, f->planes);
  }
  long long j_split = partition_cost(t, mi_r, mi_c, 16, 1);
  for (int k = 0; k < 4; k++) {
    const int xx = x + (k & 1) * 8, yy = y + (k >> 1) * 8;
    if (xx >= W || yy >= H) continue;
    j_split += encode_tree8(t, sh, xx, yy);
  }
  if (can_none && j_none <= j_split) { commit_block(t, sh.held16, x, y, 16); return j_none; }
  return j_split;
}

}  // namespace

// One CTA per tile.  Partition sizes > 16 are always split (partition_max <= 16 in this kernel).
__global__ void __launch_bounds__(kThreads) b200_rdo_tile_kernel(const B200RdoJob *jobs, const B200TileRef *tiles) {
  __shared__ Shared sh;
  const B200TileRef tr = tiles[blockIdx.x];
  const B200RdoJob &job = jobs[tr.job];
  TileCtx t;
  t.f = &job.frame; t.rt = job.rate;
  for (int p = 0; p < 3; p++) { t.src[p] = job.src[p]; t.rec[p] = job.rec[p]; t.eob[p] = job.eob[p]; t.coef[p] = job.coef[p]; }
  t.mi = job.mi; t.decoded = job.decoded; t.stride = job.frame.mi_cols * 4;
  const B200Frame *f = t.f;
  t.tr0 = f->tile_row_start_sb[tr.tile_row] * 16;
  t.tr1 = min(f->tile_row_start_sb[tr.tile_row + 1] * 16, f->mi_rows);
  t.tc0 = f->tile_col_start_sb[tr.tile_col] * 16;
  t.tc1 = min(f->tile_col_start_sb[tr.tile_col + 1] * 16, f->mi_cols);
  for (int i = threadIdx.x; i < 336; i += kThreads) { sh.scan[i] = dv_scan[i]; sh.iscan[i] = dv_iscan[i]; }
  __syncthreads();
  for (int r = t.tr0; r < t.tr1; r += 16)
    for (int c = t.tc0; c < t.tc1; c += 16)
      // 64x64 -> 32x32 -> 16x16 in z-order (always split above 16x16)
      for (int q = 0; q < 16; q++) {
        const int bx = ((q & 1) | ((q >> 1) & 2)) * 16, by = (((q >> 1) & 1) | ((q >> 2) & 2)) * 16;
        encode_tree16(t, sh, c * 4 + bx, r * 4 + by);
      }
}

#ifndef B200_EMU
extern "C" int b200_rdo_launch(const B200RdoJob *d_jobs, const B200TileRef *d_tiles, int n_tiles, void *stream) {
  if (n_tiles <= 0) return 0;
  b200_rdo_tile_kernel<<<n_tiles, kThreads, 0, (cudaStream_t)stream>>>(d_jobs, d_tiles);
  return (int)cudaGetLastError();
}
#else
static const B200RdoJob *g_emu_jobs;
static const B200TileRef *g_emu_tiles;
static void emu_entry(void *) { b200_rdo_tile_kernel(g_emu_jobs, g_emu_tiles); }
extern "C" int b200_rdo_launch_emu(const B200RdoJob *jobs, const B200TileRef *tiles, int n_tiles) {
  g_emu_jobs = jobs; g_emu_tiles = tiles;
  emu::launch(emu_entry, nullptr, (unsigned)n_tiles, kThreads);
  return 0;
}
#endif
