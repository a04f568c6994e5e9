// AV1 intra rate-distortion search on the device, in three stages:
//
//   A. ANALYSIS (open loop, embarrassingly parallel).  Every block of every level (4x4, 8x8, 16x16, 32x32 when the
//      partition range reaches it) of every image is analysed independently, predicting from SOURCE neighbours, from
//      the largest level down so that a level can be pruned by its parent (b200_pruned_by_parent, records.h):
//      luma SATD ranking of the 13 intra modes + angle refinement (b200_satd_direct_kernel for the never-pruned top
//      level, b200_satd_kernel with per-CTA compaction of the surviving blocks below it: one thread per block), then
//      full trials of the top-2 luma modes x transform types and of the chroma candidates {luma mode, DC, CFL}
//      (b200_analyze_kernel: one warp per block, lane-batched candidates: lane group g = lane / N owns one candidate,
//      lane l = lane % N owns column l / row l of its transform; butterflies in registers, transpose through shared
//      memory, quantiser / rate / dequantiser on the row, SSE by shuffles; the top level runs its warps in lock-step
//      CTAs for instruction-cache locality).
//   B. PARTITION DECISION (b200_decide_kernel: one thread per root region, NONE vs SPLIT recursively).
//   C. RECONSTRUCTION (closed loop, b200_recon_kernel: one CTA per TILE, warps on a superblock-row wavefront): the only
//      stage that needs reconstructed neighbours; one trial per decided block and plane, Y/U/V batched over the lane
//      groups, CFL chroma after its luma.
//
// Only stage C is a serial chain (a tile's blocks depend on their left/top reconstruction); stages A/B carry >95 % of
// the arithmetic and scale with the number of blocks in the batch.
//
// Replaces rav1e's encode_tile / rdo_mode_decision / encode_tx_block [EXT: crates.io rav1e 0.8.1], which the
// reference reaches only through encode_to_av1 (/root/reference/ravif/src/av1encoder.rs:749-771).  rav1e searches
// closed-loop inside one serial chain per tile; the open-loop split is this implementation's GPU-first choice
// (encoder freedom; the bitstream and the reconstruction stay exact).  The bit-exact CPU restatement of THIS file is
// oracle/av1_encode.c; tests/test_gpu_parity.py compares every stage.
//
// No tensor cores on purpose: the work is 4..32-point integer butterflies with normative rounding at every stage
// (BASELINE.json north_star).
#ifdef B200_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#ifndef B200_EMU
#include <mutex>
#endif

#include "records.h"
#include "rdo.h"

#define AV1_TABLE(type, name, dims) __constant__ const type name dims
#include "av1_tables_dev.h"
#undef AV1_TABLE

#define TXFM_FN __device__ __forceinline__
#include "txfm1d_gen.h"

namespace {

constexpr int kThreads = 32;
constexpr int EOFF = 16;          // edge arrays: index -1 .. 2n-1 stored at EOFF-1 ..
constexpr int kEdge = EOFF + 34;  // n <= 16 (per-thread edge arrays of the SATD kernel use EOFF + 66 at the 32x32 level)
constexpr unsigned FULL = 0xffffffffu;
#ifdef B200_EMU
#define B200_PREFETCH_L2(p) ((void)(p))
#else
#define B200_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif
constexpr long long kInf = 0x7fffffffffffffffll;

// inverse scan tables (4x4 | 8x8 | 16x16) in GLOBAL memory: indexed per lane (divergent), which the constant
// cache would serialise; __ldg keeps them L1 resident.
__device__ const uint8_t dv_iscan_g[336] = {
#include "iscan_init.inc"
};
__device__ const uint16_t dv_iscan32_g[1024] = {
#include "iscan32_init.inc"
};
__constant__ const uint8_t dv_sm_weights[4 + 8 + 16 + 32] = {255, 149, 85, 64, 255, 197, 146, 105, 73, 50, 37, 32,
                                                             255, 225, 196, 170, 145, 123, 102, 84, 68, 54, 43, 33, 26, 20, 17, 16,
                                                             255, 240, 225, 210, 196, 182, 169, 157, 145, 133, 122, 111, 101, 92, 83, 74,
                                                             66, 59, 52, 45, 39, 34, 29, 25, 21, 17, 14, 12, 10, 9, 8, 8};
__constant__ const uint8_t dv_tx_types[5] = {TX_DCT_DCT, TX_ADST_DCT, TX_DCT_ADST, TX_ADST_ADST, TX_IDTX};

// Per-warp working set.  NMAX = largest block size the kernel instance handles (16: 8.7 KB, 32: 22.9 KB).
template <int NMAX>
struct SharedT {
  static constexpr int kMaxN = NMAX;
  static constexpr int kPredStride = NMAX * NMAX;          // one prediction slot (luma: two ranked modes, two slots)
  uint16_t src[3][NMAX * NMAX];
  uint16_t above[3][EOFF + 2 * NMAX + 2], left[3][EOFF + 2 * NMAX + 2];
  uint16_t ef[2][EOFF + 2 * NMAX + 8];   // intra edge filter: filtered / upsampled private copy of one plane's edges
  uint16_t pred[2 * NMAX * NMAX];   // luma: predictions of the two ranked modes; chroma: up to G (mode rank, plane) predictions
  int32_t buf[32 * (NMAX + 1)];     // G groups x N rows x (N+1): transform transpose staging
  alignas(16) uint8_t lev[NMAX == 32 ? 1296 : 800];   // per group (N+4)^2 padded |level| map (rate contexts)
  int16_t qc[NMAX == 32 ? 1024 : 512];                // per group n*n quantised levels in scan order
  uint16_t rec[NMAX == 32 ? 1024 : 512];              // per group n*n reconstruction
  int16_t ac[NMAX * NMAX];          // CFL: (luma << 3) - average of the block
  uint32_t taskD[6], taskR[6];
  uint16_t taskE[6];
};
typedef SharedT<16> Shared;

struct TileCtx {
  const B200Frame *f;
  const B200RateTables *rt;
  const uint16_t *src[3];
  uint16_t *rec[3];
  B200Mi *mi;
  uint16_t *eob[3];
  int16_t *coef[3];
  int stride;
  int tr0, tr1, tc0, tc1;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int txs_of(int n) { return n == 4 ? 0 : n == 8 ? 1 : n == 16 ? 2 : 3; }
__device__ __forceinline__ int scan_base(int n) { return n == 4 ? 0 : n == 8 ? 16 : 80; }
// position -> scan index (tables in global memory, L1 resident: the index diverges per lane)
template <int N> __device__ __forceinline__ int iscan_at(int pos) {
  if (N == 32) return __ldg(&dv_iscan32_g[pos]);
  return __ldg(&dv_iscan_g[scan_base(N) + pos]);
}
__device__ __forceinline__ int ilog2_dev(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

__device__ __noinline__ long long warp_sum_ll(long long v) {
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
template <int N> __device__ __forceinline__ uint32_t group_sum(uint32_t v) {
#pragma unroll
  for (int m = N / 2; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
template <int N> __device__ __forceinline__ int group_max(int v) {
#pragma unroll
  for (int m = N / 2; m > 0; m >>= 1) v = max(v, __shfl_xor_sync(FULL, v, m));
  return v;
}
// lexicographic minimum of (J, idx) over the warp: J >= 0 and idx < 32 are packed into one 64-bit key
// (J < 2^58: SSE*256 of a 16x16 10-bit block is < 2^37); the 64-bit minimum is two hardware warp reductions (REDUX)
__device__ __forceinline__ void warp_argmin(long long &J, int &idx) {
  const unsigned long long key = J == kInf ? ~0ull : (((unsigned long long)J << 5) | (unsigned)(idx & 31));
  const unsigned hi = (unsigned)(key >> 32);
  const unsigned mh = __reduce_min_sync(FULL, hi);
  const unsigned ml = __reduce_min_sync(FULL, hi == mh ? (unsigned)key : 0xffffffffu);
  const unsigned long long best = ((unsigned long long)mh << 32) | ml;
  if (best == ~0ull) { J = kInf; idx = 1 << 20; }
  else { J = (long long)(best >> 5); idx = (int)(best & 31); }
}

// ---- edges (AV1 spec 7.11.2, intra edge filter disabled) from plane `r` (source in the analysis, reconstruction in
// the final pass) into sh.above/left[slot]; top-right / bottom-left availability = static coding order predicate
// avail: bit0 have_above, bit1 have_left, bit2 have_above_right, bit3 have_below_left (block_avail(), once per block)
__device__ __noinline__ int block_avail(const TileCtx &t, int x, int y, int n) {
  const int mi_r = y >> 2, mi_c = x >> 2, n4 = n >> 2;
  const int ha = mi_r > t.tr0, hl = mi_c > t.tc0;
  const int ar = ha && b200_unit_before(mi_r - 1, mi_c + n4, mi_r, mi_c, t.tr0, t.tr1, t.tc0, t.tc1);
  const int bl = hl && b200_unit_before(mi_r + n4, mi_c - 1, mi_r, mi_c, t.tr0, t.tr1, t.tc0, t.tc1);
  return ha | (hl << 1) | (ar << 2) | (bl << 3);
}
template <class SH>
__device__ __noinline__ void build_edges(const TileCtx &t, SH &sh, const uint16_t *r, int slot, int x, int y, int n, int avail) {
  const B200Frame *f = t.f;
  const int st = t.stride, lane = lane_id();
  const int ha = avail & 1, hl = (avail >> 1) & 1, have_ar = (avail >> 2) & 1, have_bl = (avail >> 3) & 1;
  const int max_x = f->mi_cols * 4 - 1, max_y = f->mi_rows * 4 - 1;
  const int base = 1 << (f->bit_depth - 1);
  uint16_t *A = sh.above[slot] + EOFF, *L = sh.left[slot] + EOFF;
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
  }
}

// ---- intra predictors (AV1 spec 7.11.2.x): value of pixel (i, j) of an n x n block ------------------
__device__ __noinline__ int dc_value(int n, const uint16_t *A, const uint16_t *L, int ha, int hl, int bd) {
  if (!ha && !hl) return 1 << (bd - 1);
  const int lg = 31 - __clz(n);          // n is a power of two: the spec's divisions are shifts
  int s = 0;
#pragma unroll 1
  for (int i = 0; i < n; i++) s += (ha ? A[i] : 0) + (hl ? L[i] : 0);
  return (ha && hl) ? (s + n) >> (lg + 1) : (s + (n >> 1)) >> lg;
}
// the same value computed by the whole warp (edges in shared memory): lane i adds sample i (n <= 32), one REDUX
__device__ __noinline__ int dc_value_warp(int n, const uint16_t *A, const uint16_t *L, int ha, int hl, int bd) {
  if (!ha && !hl) return 1 << (bd - 1);
  const int lg = 31 - __clz(n), i = lane_id();
  const unsigned s = __reduce_add_sync(FULL, i < n ? (unsigned)((ha ? A[i] : 0) + (hl ? L[i] : 0)) : 0u);
  return (ha && hl) ? (int)(s + n) >> (lg + 1) : (int)(s + (n >> 1)) >> lg;
}
// A, L point at index 0 of the edge arrays (index -1 valid).  dc: precomputed DC value (used for DC_PRED only).
__device__ __forceinline__ int pred_px(int mode, int delta, int n, int i, int j, const uint16_t *A, const uint16_t *L, int dc) {
  if (mode == DC_PRED) return dc;
  if (mode == PAETH_PRED) {
    const int tl = A[-1], b = A[j] + L[i] - tl;
    const int pl = abs(b - L[i]), pt = abs(b - A[j]), ptl = abs(b - tl);
    return (pl <= pt && pl <= ptl) ? L[i] : (pt <= ptl ? A[j] : tl);
  }
  if (mode >= SMOOTH_PRED) {   // SMOOTH, SMOOTH_V, SMOOTH_H share the two half sums
    const uint8_t *w = dv_sm_weights + (n == 4 ? 0 : n == 8 ? 4 : n == 16 ? 12 : 28);
    const int v = w[i] * A[j] + (256 - w[i]) * L[n - 1], h = w[j] * L[i] + (256 - w[j]) * A[n - 1];
    return mode == SMOOTH_PRED ? (v + h + 256) >> 9 : (mode == SMOOTH_V_PRED ? (v + 128) >> 8 : (h + 128) >> 8);
  }
  const int p_angle = dv_mode_to_angle[mode] + 3 * delta;
  if (p_angle == 90) return A[j];
  if (p_angle == 180) return L[i];
  if (p_angle > 90 && p_angle < 180) {
    const int dx = dv_dr_intra_derivative[180 - p_angle], dy = dv_dr_intra_derivative[p_angle - 90];
    int idx = (j << 6) - (i + 1) * dx, base = idx >> 6;
    const uint16_t *e = A;
    if (base < -1) { idx = (i << 6) - (j + 1) * dy; base = idx >> 6; e = L; }
    const int shift = (idx >> 1) & 31;
    return (e[base] * (32 - shift) + e[base + 1] * shift + 16) >> 5;
  }
  // zones 1 (< 90, above row) and 3 (> 180, left column) are the same walk with the roles of i / j swapped
  const bool z1 = p_angle < 90;
  const uint16_t *e = z1 ? A : L;
  const int d = dv_dr_intra_derivative[z1 ? p_angle : 270 - p_angle], max_base = 2 * n - 1;
  const int idx = ((z1 ? i : j) + 1) * d, base = (idx >> 6) + (z1 ? j : i), shift = (idx >> 1) & 31;
  return base < max_base ? (e[base] * (32 - shift) + e[base + 1] * shift + 16) >> 5 : e[max_base];
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
    else if (N == 16) out[i] = round_shift64((long long)in[i] * 2 * 5793, 12);
    else out[i] = in[i] * 4;
  }
}
// HBD = false (8-bit video): every half-butterfly w0*a + w1*b fits int32 (DESIGN.md "int32 butterflies"), so the
// multiply-accumulate is a single IMAD; HBD = true (10-bit) keeps the int64 form of the specification.
template <int N, bool HBD>
__device__ __forceinline__ void fwd1d(int kind, const int32_t *in, int32_t *out, int bit) {
  const int32_t *c = dv_cospi_arr[bit - 10];
  if (kind == K_IDT) { identity_dev<N>(in, out); return; }
  if (N == 32) { gen_fdct32(in, out, c, bit); return; }      // 32x32 is DCT (or identity) only
  if (kind == K_DCT) {
    if (HBD) { if (N == 4) gen_fdct4(in, out, c, bit); else if (N == 8) gen_fdct8(in, out, c, bit); else gen_fdct16(in, out, c, bit); }
    else { if (N == 4) gen_fdct4_w32(in, out, c, bit); else if (N == 8) gen_fdct8_w32(in, out, c, bit); else gen_fdct16_w32(in, out, c, bit); }
  } else {
    if (N == 4) fadst4_dev(in, out, bit);
    else if (HBD) { if (N == 8) gen_fadst8(in, out, c, bit); else gen_fadst16(in, out, c, bit); }
    else { if (N == 8) gen_fadst8_w32(in, out, c, bit); else gen_fadst16_w32(in, out, c, bit); }
  }
}
template <int N, bool HBD>
__device__ __forceinline__ void inv1d(int kind, const int32_t *in, int32_t *out) {
  const int bit = 12;
  const int32_t *c = dv_cospi_arr[bit - 10];
  if (kind == K_IDT) { identity_dev<N>(in, out); return; }
  if (N == 32) { gen_idct32(in, out, c, bit); return; }
  if (kind == K_DCT) {
    if (HBD) { if (N == 4) gen_idct4(in, out, c, bit); else if (N == 8) gen_idct8(in, out, c, bit); else gen_idct16(in, out, c, bit); }
    else { if (N == 4) gen_idct4_w32(in, out, c, bit); else if (N == 8) gen_idct8_w32(in, out, c, bit); else gen_idct16_w32(in, out, c, bit); }
  } else {
    if (N == 4) iadst4_dev(in, out, bit);
    else if (HBD) { if (N == 8) gen_iadst8(in, out, c, bit); else gen_iadst16(in, out, c, bit); }
    else { if (N == 8) gen_iadst8_w32(in, out, c, bit); else gen_iadst16_w32(in, out, c, bit); }
  }
}
__device__ __forceinline__ int32_t rshift_round(int32_t v, int bit) {   // bit > 0: rounding right shift; <= 0: left shift
  return bit <= 0 ? (int32_t)((uint32_t)v << (-bit)) : (int32_t)(((long long)v + (1ll << (bit - 1))) >> bit);
}
__device__ __forceinline__ int32_t clamp_bits(int32_t v, int bits) {
  const int32_t mx = (1 << (bits - 1)) - 1, mn = -(1 << (bits - 1));
  return v < mn ? mn : (v > mx ? mx : v);
}

// ---- one lane batch: G = 32/N candidates of an N x N block evaluated in lock-step -------------------
// Lane group g = lane / N, lane-in-group l = lane % N.  plane / src / pred / tx_type / y_mode may differ per group
// (identical for the lanes of one group).  Results (eob, rate, sse) are returned to every lane of the group; the
// group's quantised levels (scan order) and reconstruction stay in sh.qc / sh.rec at offset g*N*N.
// Mirrors oracle/av1_encode.c trial_plane(): forward transform (libaom av1_fwd_txfm2d design), quantize(),
// coef_rate(), AV1 spec 7.13.3 inverse transform + reconstruction, SSE.
// Code size matters (the instruction cache is the first bottleneck of this kernel): the function is NOT inlined,
// each 1-D transform kernel appears once (two-pass loops), and the per-coefficient loops are rolled.
struct TrialOut { int eob; uint32_t rate, sse; };

// RATE = false (reconstruction stage): the coefficient-rate model is skipped, TrialOut.rate is 0.
template <int N, bool HBD, class SH, bool RATE = true>
__device__ __noinline__ TrialOut trial_batch(const TileCtx &t, SH &sh, int plane, const uint16_t *src, const uint16_t *pred,
                                             int tx_type, int y_mode) {
  static_assert(N <= SH::kMaxN, "working set too small for this transform size");
  constexpr int S = N + 1, N2 = N * N, LS = N + 4;
  constexpr int shift1 = N == 4 ? 0 : (N == 8 ? 1 : (N == 16 ? 2 : 4));   // -fwd_shift[1] (libaom fwd_txfm_shift_ls)
  constexpr int row_shift = N == 4 ? 0 : (N == 8 ? 1 : 2);                // inverse: shift after the row pass
  constexpr int qshift = N == 32 ? 1 : 0;                                 // 32x32: coefficients are coded at half scale
  const B200Frame *f = t.f;
  const B200RateTables *rt = t.rt;
  const int lane = lane_id(), g = lane / N, l = lane % N;
  int32_t *buf = sh.buf + g * N * S;
  uint8_t *lev = sh.lev + g * LS * LS;
  int16_t *qc = sh.qc + g * N2;
  uint16_t *rec = sh.rec + g * N2;
  int vert, horz;
  tx_kernels(tx_type, vert, horz);
  const int bd = f->bit_depth;

  // ---- forward: pass 0 = column l (residual << 2 -> column transform -> buf), pass 1 = row l (buf -> row transform -> buf)
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    int32_t x[N], o[N];
    if (pass == 0) {
#pragma unroll
      for (int r = 0; r < N; r++) x[r] = ((int)src[r * N + l] - (int)pred[r * N + l]) * 4;
    } else {
#pragma unroll
      for (int c = 0; c < N; c++) x[c] = buf[l * S + c];
    }
    fwd1d<N, HBD>(pass == 0 ? vert : horz, x, o, (N == 32 || (pass == 1 && N == 16)) ? 12 : 13);   // fwd_cos_bit_col / _row
    if (pass == 0) {
#pragma unroll
      for (int r = 0; r < N; r++) buf[r * S + l] = rshift_round(o[r], shift1);
      __syncwarp();
    } else {
#pragma unroll
      for (int c = 0; c < N; c++) buf[l * S + c] = o[c];
    }
  }
  // ---- quantise / dequantise row l in place (oracle quantize()); exact division by reciprocal multiplication
  int eob = 0;
  {
    const uint32_t dc_q = (uint32_t)f->dc_q, ac_q = (uint32_t)f->ac_q;
    const uint32_t dc_rcp = 0xffffffffu / dc_q, ac_rcp = 0xffffffffu / ac_q;
#pragma unroll 1
    for (int c = 0; c < N; c++) {
      const int pos = l * N + c;
      const uint32_t q = pos == 0 ? dc_q : ac_q, rcp = pos == 0 ? dc_rcp : ac_rcp;
      const uint32_t off = pos == 0 ? (q * 109) >> 8 : (q * 98) >> 8;
      const int32_t cf = buf[l * S + c];
      const uint32_t a = ((uint32_t)(cf < 0 ? -cf : cf) << qshift) + off;
      uint32_t lv = __umulhi(a, rcp);           // floor(a / q) or one/two less
      uint32_t rem = a - lv * q;
      if (rem >= q) { lv++; rem -= q; }
      if (rem >= q) lv++;
      if (lv > 32767) lv = 32767;
      const int32_t dqa = (int32_t)((lv * q) & 0xFFFFFF) >> qshift;
      const int si = iscan_at<N>(pos);
      qc[si] = (int16_t)(cf < 0 ? -(int)lv : (int)lv);
      buf[l * S + c] = cf < 0 ? -dqa : dqa;
      lev[l * LS + c] = (uint8_t)(lv > 127 ? 127 : lv);
      if (lv) eob = max(eob, si + 1);
    }
  }
  eob = group_max<N>(eob);
  __syncwarp();
  // ---- rate (oracle coef_rate())
  const int txs = txs_of(N), ptype = plane > 0;
  const int skip_ctx = plane == 0 ? 0 : 7;
  uint32_t rate = 0;
  if (!RATE) {
  } else if (eob == 0) {
    if (l == 0) rate = rt->txb_skip[txs][skip_ctx][1];
  } else {
    if (l == 0) {
      rate = rt->txb_skip[txs][skip_ctx][0];
      if (plane == 0 && N <= 16) rate += rt->intra_tx[(f->reduced_tx_set || N == 16) ? 1 : 0][txs][y_mode][tx_type];
      int tg = 0;
      while (tg < 11 && (tg == 0 ? 1 : (tg == 1 ? 2 : (1 << (tg - 1)) + 1)) <= eob) tg++;   // group_start[tg + 1] <= eob
      const int gstart = tg < 3 ? tg : (1 << (tg - 2)) + 1;                                 // group_start[tg]
      rate += rt->eob_pt[txs][ptype][tg - 1];
      const int nb = tg < 3 ? 0 : tg - 2;
      if (nb > 0) {
        const int extra = eob - gstart;
        rate += rt->eob_extra[txs][ptype][tg - 3][(extra >> (nb - 1)) & 1];
        rate += 256u * (uint32_t)(nb - 1);
      }
    }
#pragma unroll 1
    for (int c = 0; c < N; c++) {
      const int pos = l * N + c;
      const int si = iscan_at<N>(pos);
      if (si >= eob) continue;
      const int level = abs((int)qc[si]);
      const uint8_t *L = lev + l * LS + c;
      if (si == eob - 1) {
        const int c4 = si == 0 ? 0 : (si <= N2 / 8 ? 1 : (si <= N2 / 4 ? 2 : 3));
        rate += rt->coeff_base_eob[txs][ptype][c4][min(level, 3) - 1];
      } else {
        const int mag = min((int)L[1], 3) + min((int)L[LS], 3) + min((int)L[LS + 1], 3) + min((int)L[2], 3) + min((int)L[2 * LS], 3);
        int ctx = min((mag + 1) >> 1, 4);
        if (l == 0 && c == 0) ctx = 0;
        else ctx += (l + c < 2) ? 1 : ((l + c < 4) ? 6 : 21);
        rate += rt->coeff_base[txs][ptype][ctx][min(level, 3)];
      }
      if (level > 2) {
        int mag = min((int)L[1], 15) + min((int)L[LS], 15) + min((int)L[LS + 1], 15);
        mag = min((mag + 1) >> 1, 6);
        const int bctx = (l == 0 && c == 0) ? mag : ((l < 2 && c < 2) ? mag + 7 : mag + 14);
        rate += rt->coeff_br[txs][ptype][bctx][min(level - 3, 12)];
        if (level > 14) rate += 256u * (uint32_t)(2 * (32 - __clz(level - 14)) - 1);   // golomb(level - 15)
      }
      if (level) rate += (si == 0) ? rt->dc_sign[ptype][0][qc[si] < 0] : 256u;
    }
  }
  if (RATE) rate = group_sum<N>(rate);
  // ---- inverse (AV1 spec 7.13.3): pass 0 = row l (buf -> buf), pass 1 = column l -> reconstruction + SSE
  uint32_t sse = 0;
  if (eob) {
    const int col_clamp = bd + 6 > 16 ? bd + 6 : 16;
    const int mx = (1 << bd) - 1;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      int32_t x[N], o[N];
      if (pass == 0) {
#pragma unroll
        for (int c = 0; c < N; c++) x[c] = clamp_bits(buf[l * S + c], bd + 8);
      } else {
#pragma unroll
        for (int r = 0; r < N; r++) x[r] = clamp_bits(buf[r * S + l], col_clamp);
      }
      inv1d<N, HBD>(pass == 0 ? horz : vert, x, o);
      if (pass == 0) {
#pragma unroll
        for (int c = 0; c < N; c++) buf[l * S + c] = rshift_round(o[c], row_shift);
      } else {
#pragma unroll
        for (int r = 0; r < N; r++) {
          int v = (int)pred[r * N + l] + rshift_round(o[r], 4);
          v = v < 0 ? 0 : (v > mx ? mx : v);
          rec[r * N + l] = (uint16_t)v;
          const int d = (int)src[r * N + l] - v;
          sse += (uint32_t)(d * d);
        }
      }
      __syncwarp();     // groups with eob == 0 skip this loop: see the matching barriers below
    }
  } else {
#pragma unroll 1
    for (int r = 0; r < N; r++) {
      const int v = pred[r * N + l];
      rec[r * N + l] = (uint16_t)v;
      const int d = (int)src[r * N + l] - v;
      sse += (uint32_t)(d * d);
    }
    __syncwarp();
    __syncwarp();
  }
  sse = group_sum<N>(sse);
  __syncwarp();
  TrialOut o;
  o.eob = eob; o.rate = rate; o.sse = sse;
  return o;
}

template <bool HBD, class SH>
__device__ __forceinline__ TrialOut trial_dispatch(int n, const TileCtx &t, SH &sh, int plane, const uint16_t *src,
                                                   const uint16_t *pred, int tx_type, int y_mode) {
  // only the reconstruction stage dispatches on a run-time size, and it does not need the rate
  if (n == 4) return trial_batch<4, HBD, SH, false>(t, sh, plane, src, pred, tx_type, y_mode);
  if (n == 8) return trial_batch<8, HBD, SH, false>(t, sh, plane, src, pred, tx_type, y_mode);
  if (SH::kMaxN >= 32 && n == 32) return trial_batch<(SH::kMaxN >= 32 ? 32 : 16), HBD, SH, false>(t, sh, plane, src, pred, tx_type, y_mode);
  return trial_batch<16, HBD, SH, false>(t, sh, plane, src, pred, tx_type, y_mode);
}

__device__ __forceinline__ long long rd_cost(const B200Frame *f, unsigned long long sse, uint32_t rate) {
  return (long long)(sse << 8) + (((long long)f->lambda_q8 * (long long)rate) >> 8);
}
__device__ __forceinline__ int chroma_tx_type(int uv_mode) { return uv_mode == UV_CFL_PRED ? 0 : dv_intra_mode_to_tx_type[uv_mode]; }
__device__ __forceinline__ uint32_t mode_rate(const B200RateTables *rt, const uint16_t *cost_row, int m, int n, int delta) {
  uint32_t r = cost_row[m];
  if (n >= 8 && m >= V_PRED && m <= D67_PRED) r += rt->angle_delta[m - V_PRED][delta + 3];
  return r;
}

// ---- intra edge filter / upsampling (AV1 spec 7.11.2.7 - 7.11.2.12; oracle predict_directional_ief), warp-cooperative
__device__ __forceinline__ int ief_strength(int blk_wh, int filter_type, int delta) {
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
__device__ __forceinline__ int ief_use_upsample(int blk_wh, int filter_type, int delta) {
  const int d = abs(delta);
  if (d <= 0 || d >= 40) return 0;
  return filter_type == 0 ? blk_wh <= 16 : blk_wh <= 8;
}
// buf: index -1 .. sz-2 filtered in place (entries 0 .. sz-2 change), sz <= 65
__device__ __forceinline__ void ief_filter_edge(uint16_t *buf, int sz, int strength) {
  if (strength) {
    const int k0 = strength == 3 ? 2 : 0, k1 = strength == 1 ? 4 : (strength == 2 ? 5 : 4), k2 = strength == 1 ? 8 : (strength == 2 ? 6 : 4);
    uint16_t o[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const int i = 1 + lane_id() + 32 * r;
      int sum = 0;
      if (i < sz) {
        const int e0 = buf[max(i - 2, 0) - 1], e1 = buf[i - 2], e2 = buf[i - 1], e3 = buf[min(i + 1, sz - 1) - 1], e4 = buf[min(i + 2, sz - 1) - 1];
        sum = k0 * (e0 + e4) + k1 * (e1 + e3) + k2 * e2;
      }
      o[r] = (uint16_t)((sum + 8) >> 4);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; r++) { const int i = 1 + lane_id() + 32 * r; if (i < sz) buf[i - 1] = o[r]; }
  }
  __syncwarp();
}
// buf: index -2 .. 2 * num_px - 2 after the call, num_px <= 16
__device__ __forceinline__ void ief_upsample_edge(uint16_t *buf, int num_px, int bd) {
  const int i = lane_id();
  int v = 0, m = 0;
  const int corner = buf[-1];
  if (i < num_px) {
    const int d0 = i == 0 ? corner : buf[i - 2], d1 = buf[i - 1], d2 = buf[i], d3 = buf[min(i + 1, num_px - 1)];
    v = (-d0 + 9 * d1 + 9 * d2 - d3 + 8) >> 4;
    const int mx = (1 << bd) - 1;
    v = v < 0 ? 0 : (v > mx ? mx : v);
    m = d2;
  }
  __syncwarp();
  if (i < num_px) { buf[2 * i - 1] = (uint16_t)v; buf[2 * i] = (uint16_t)m; }
  if (i == 0) buf[-2] = (uint16_t)corner;
  __syncwarp();
}
__device__ __forceinline__ int pred_dir_px(int p_angle, int n, int i, int j, const uint16_t *A, const uint16_t *L, int up_a, int up_l) {
  if (p_angle == 90) return A[j];
  if (p_angle == 180) return L[i];
  if (p_angle < 90) {
    const int dx = dv_dr_intra_derivative[p_angle], max_base = (2 * n - 1) << up_a;
    const int idx = (i + 1) * dx, base = (idx >> (6 - up_a)) + (j << up_a), shift = ((idx << up_a) >> 1) & 31;
    return base < max_base ? (A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5 : A[max_base];
  }
  if (p_angle < 180) {
    const int dx = dv_dr_intra_derivative[180 - p_angle], dy = dv_dr_intra_derivative[p_angle - 90];
    int idx = (j << 6) - (i + 1) * dx, base = idx >> (6 - up_a);
    if (base >= -(1 << up_a)) {
      const int shift = ((idx << up_a) >> 1) & 31;
      return (A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5;
    }
    idx = (i << 6) - (j + 1) * dy; base = idx >> (6 - up_l);
    const int shift = ((idx << up_l) >> 1) & 31;
    return (L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5;
  }
  const int dy = dv_dr_intra_derivative[270 - p_angle];
  const int idx = (j + 1) * dy, base = (idx >> (6 - up_l)) + (i << up_l), shift = ((idx << up_l) >> 1) & 31;
  return (L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5;
}
// directional mode with the edge filter on: ef[0] / ef[1] receive the private copies of the edges
__device__ __noinline__ void predict_block_ief(int mode, int delta, int lg, const uint16_t *above, const uint16_t *left, int ha, int hl,
                                               int bd, int ftype, uint16_t *ef_a, uint16_t *ef_l, uint16_t *dst) {
  const int n = 1 << lg, p_angle = dv_mode_to_angle[mode] + 3 * delta;
  uint16_t *A = ef_a + EOFF, *L = ef_l + EOFF;
#pragma unroll 1
  for (int i = lane_id() - 1; i < 2 * n; i += 32) { A[i] = above[EOFF + i]; L[i] = left[EOFF + i]; }
  __syncwarp();
  if (p_angle != 90 && p_angle != 180) {
    if (p_angle > 90 && p_angle < 180 && n >= 16) {
      const int c = (L[0] * 5 + A[-1] * 6 + A[0] * 5 + 8) >> 4;
      __syncwarp();
      if (lane_id() == 0) { A[-1] = (uint16_t)c; L[-1] = (uint16_t)c; }
      __syncwarp();
    }
    if (ha) ief_filter_edge(A, n + (p_angle < 90 ? n : 0) + 1, ief_strength(2 * n, ftype, p_angle - 90));
    if (hl) ief_filter_edge(L, n + (p_angle > 180 ? n : 0) + 1, ief_strength(2 * n, ftype, p_angle - 180));
  }
  const int up_a = ief_use_upsample(2 * n, ftype, p_angle - 90), up_l = ief_use_upsample(2 * n, ftype, p_angle - 180);
  if (up_a) ief_upsample_edge(A, n + (p_angle < 90 ? n : 0), bd);
  if (up_l) ief_upsample_edge(L, n + (p_angle > 180 ? n : 0), bd);
#pragma unroll 1
  for (int p = lane_id(); p < n * n; p += 32) dst[p] = (uint16_t)pred_dir_px(p_angle, n, p >> lg, p & (n - 1), A, L, up_a, up_l);
  __syncwarp();
}

// whole-block prediction into shared memory, pixels spread over the 32 lanes (the only place pred_px is expanded)
__device__ __noinline__ void predict_block(int mode, int delta, int lg, const uint16_t *above, const uint16_t *left, int ha, int hl, int bd,
                                           uint16_t *dst) {
  const uint16_t *A = above + EOFF, *L = left + EOFF;
  const int n = 1 << lg;
  const int dc = mode == DC_PRED ? dc_value_warp(n, A, L, ha, hl, bd) : 0;
#pragma unroll 1
  for (int p = lane_id(); p < n * n; p += 32) dst[p] = (uint16_t)pred_px(mode, delta, n, p >> lg, p & (n - 1), A, L, dc);
}

// ftype: -1 = intra edge filter off, 0 / 1 = on with that filter type
template <class SH>
__device__ __forceinline__ void predict_any(SH &sh, int ftype, int mode, int delta, int lg, const uint16_t *above, const uint16_t *left, int ha,
                                            int hl, int bd, uint16_t *dst) {
  if (ftype >= 0 && mode >= V_PRED && mode <= D67_PRED) predict_block_ief(mode, delta, lg, above, left, ha, hl, bd, ftype, sh.ef[0], sh.ef[1], dst);
  else predict_block(mode, delta, lg, above, left, ha, hl, bd, dst);
}

// ---- chroma from luma (oracle cfl_ac / cfl_pick_alpha / cfl_predict), pixels spread over the 32 lanes
__device__ __noinline__ void cfl_build_ac(const uint16_t *luma, int lg, int16_t *ac) {
  const int N2 = 1 << (2 * lg);
  long long sum = 0;
#pragma unroll 1
  for (int p = lane_id(); p < N2; p += 32) sum += luma[p] << 3;
  const int avg = (int)((warp_sum_ll(sum) + (N2 >> 1)) >> (2 * lg));
#pragma unroll 1
  for (int p = lane_id(); p < N2; p += 32) ac[p] = (int16_t)((luma[p] << 3) - avg);
  __syncwarp();
}
__device__ __noinline__ int cfl_pick_alpha(const int16_t *ac, const uint16_t *src, int N2, int dc) {
  long long num = 0, den = 0;
#pragma unroll 1
  for (int p = lane_id(); p < N2; p += 32) {
    const int a = ac[p];
    num += (long long)a * ((int)src[p] - dc);
    den += (long long)a * a;
  }
  num = warp_sum_ll(num); den = warp_sum_ll(den);
  if (den == 0) return 0;
  const unsigned long long mag = (((unsigned long long)(num < 0 ? -num : num) << 6) + (unsigned long long)(den >> 1)) / (unsigned long long)den;
  const int a = mag > 16 ? 16 : (int)mag;
  return num < 0 ? -a : a;
}
__device__ __noinline__ void cfl_predict_block(const int16_t *ac, int N2, int dc, int alpha, int bd, uint16_t *dst) {
  const int mx = (1 << bd) - 1;
#pragma unroll 1
  for (int p = lane_id(); p < N2; p += 32) {
    const int sc = alpha * ac[p];
    const int v = dc + (sc < 0 ? -((-sc + 32) >> 6) : ((sc + 32) >> 6));
    dst[p] = (uint16_t)(v < 0 ? 0 : (v > mx ? mx : v));
  }
}
__device__ __noinline__ uint32_t cfl_rate(const B200RateTables *rt, int au, int av) {
  const int su = au == 0 ? 0 : (au < 0 ? 1 : 2), sv = av == 0 ? 0 : (av < 0 ? 1 : 2);
  uint32_t r = rt->cfl_sign[su * 3 + sv - 1];
  if (su) r += rt->cfl_alpha[(su - 1) * 3 + sv][abs(au) - 1];
  if (sv) r += rt->cfl_alpha[(sv - 1) * 3 + su][abs(av) - 1];
  return r;
}

// top-2 of 13 mode costs held by lanes 0..12 (kInf when not allowed): lowest cost first, ties -> lowest index
__device__ __noinline__ void rank_top2(long long c, int &m0, int &m1) {
  const int lane = lane_id();
  long long j = lane < 13 ? c : kInf;
  int i = lane & 15;
  warp_argmin(j, i);
  m0 = j == kInf ? -1 : i;
  long long j2 = (lane < 13 && lane != m0) ? c : kInf;
  int i2 = lane & 15;
  warp_argmin(j2, i2);
  m1 = j2 == kInf ? -1 : i2;
}

__device__ __forceinline__ void tile_of_unit(const B200Frame *f, int mi_r, int mi_c, TileCtx &t) {
  int r = 0, c = 0;
  while (f->tile_row_start_sb[r + 1] * 16 <= mi_r) r++;
  while (f->tile_col_start_sb[c + 1] * 16 <= mi_c) c++;
  t.tr0 = f->tile_row_start_sb[r] * 16; t.tr1 = min(f->tile_row_start_sb[r + 1] * 16, f->mi_rows);
  t.tc0 = f->tile_col_start_sb[c] * 16; t.tc1 = min(f->tile_col_start_sb[c + 1] * 16, f->mi_cols);
}
__device__ __forceinline__ void ctx_from_job(const B200RdoJob &job, TileCtx &t) {
  t.f = &job.frame; t.rt = job.rate;
  for (int p = 0; p < 3; p++) { t.src[p] = job.src[p]; t.rec[p] = job.rec[p]; t.eob[p] = job.eob[p]; t.coef[p] = job.coef[p]; }
  t.mi = job.mi; t.stride = job.frame.mi_cols * 4;
}

// ---- stage A: analysis of one block by one warp (oracle orc_analyze_block, minus the SATD table) ---------------
// LG is a template parameter so that each per-level kernel contains only ITS transform size (instruction-cache
// footprint: ~30 KB for 4x4, the level that dominates the block count).
// SYNC: the warps of the CTA run their blocks in lock-step - a CTA barrier in front of every trial batch, with fixed trip
// counts - so that they fetch the same instructions at the same time: the per-block code (~60 KB at 16x16) does not
// fit the instruction cache and `stall_no_instruction` was the top stall of the free-running kernel (profiles/).
// `out` == nullptr: a padding warp that repeats a valid block only to take part in the barriers.
template <int LG, bool HBD, bool SYNC, class SH>
__device__ __forceinline__ void analyze_block(const TileCtx &t, SH &sh, int x, int y, const uint32_t *satd, B200BlockDec *out) {
  constexpr int lg = LG, n = 1 << LG, N2 = n * n, G = 32 >> LG, PS = SH::kPredStride;
  if (SYNC) __syncthreads();
  const B200Frame *f = t.f;
  const B200RateTables *rt = t.rt;
  const int lane = lane_id(), g = lane >> lg;
  const int planes = f->planes, bd = f->bit_depth;
  // ---- stage source block and SOURCE edges; clear the padded level maps
#pragma unroll 1
  for (int p = 0; p < planes; p++)
#pragma unroll 1
    for (int q = lane; q < N2; q += 32) sh.src[p][q] = t.src[p][(y + (q >> lg)) * t.stride + x + (q & (n - 1))];
#pragma unroll 1
  for (int i = lane; i < (int)sizeof(sh.lev) / 4; i += 32) reinterpret_cast<uint32_t *>(sh.lev)[i] = 0;
  const int avail = block_avail(t, x, y, n), ha = avail & 1, hl = (avail >> 1) & 1;
#pragma unroll 1
  for (int p = 0; p < planes; p++) build_edges(t, sh, t.src[p], p, x, y, n, avail);
  __syncwarp();
  // ---- ranking from the SATD table (context-0 mode rates)
  const uint16_t *ymode_cost = rt->y_mode[0][0];
  const uint32_t wy = satd[13];      // refined directional mode | (delta + 3) << 8
  const int md_y = wy & 255, dl_y = (int)(wy >> 8) - 3;
#define DELTA_Y(m) ((m) == md_y ? dl_y : 0)
#define DELTA_UV(m) ((m) == best_mode ? DELTA_Y(m) : 0)
  long long my_ycost = kInf;
  if (lane < 13 && ((f->y_mode_mask >> lane) & 1))
    my_ycost = (long long)satd[lane] * 64 + (((long long)f->lambda_satd_q8 * mode_rate(rt, ymode_cost, lane, n, DELTA_Y(lane))) >> 8);
  int m0, m1;
  rank_top2(my_ycost, m0, m1);
  // ---- luma full trials: candidate c = k * 5 + ti; lane batches pair the two modes on the same transform type
  const int tx_mask = n == 32 ? 1 : f->tx_type_mask;      // 32x32: the intra transform set is DCT only
  int first_t = -1;
  for (int k = 0; k < 5; k++) if ((tx_mask >> k) & 1) { first_t = k; break; }
  // open loop: the neighbours' modes are not decided yet, the trials assume filter type 0 (oracle)
  const int ief = f->enable_intra_edge_filter ? 0 : -1;
  predict_any(sh, ief, m0 >= 0 ? m0 : 0, DELTA_Y(m0 >= 0 ? m0 : 0), lg, sh.above[0], sh.left[0], ha, hl, bd, sh.pred);
  predict_any(sh, ief, m1 >= 0 ? m1 : (m0 >= 0 ? m0 : 0), DELTA_Y(m1 >= 0 ? m1 : (m0 >= 0 ? m0 : 0)), lg, sh.above[0], sh.left[0], ha, hl, bd, sh.pred + PS);
  __syncwarp();
  long long bestJ = kInf;
  int bestC = 1 << 20, bestEob = 0;
  bool skip_rest = false;      // every first-type trial quantised to all-zero: the other transform types are not tried
#pragma unroll 1
  for (int b = 0; b * G < 10 && (SYNC || !skip_rest); b++) {
    if (SYNC) { __syncthreads(); if (skip_rest) continue; }
    const int ci = b * G + g;
    const int ti = ci >> 1, k = ci & 1;
    const int m = k == 0 ? m0 : m1;
    bool ok = ti < 5 && m >= 0 && ((tx_mask >> ti) & 1) && (f->rdo_tx_decision || ti == first_t);
    if (!__ballot_sync(FULL, ok)) continue;
    const TrialOut r = trial_batch<n, HBD>(t, sh, 0, sh.src[0], sh.pred + (ok ? k : 0) * PS, dv_tx_types[ok ? ti : 0],
                                           ok ? m : (m0 >= 0 ? m0 : 0));
    const int gf = first_t * 2 - b * G;        // lane group of (rank 0, first type) in this batch, if it is here
    if (G >= 2 && f->rdo_tx_decision && gf >= 0 && gf < G) {
      const int e0 = __shfl_sync(FULL, r.eob, gf << lg), e1 = __shfl_sync(FULL, r.eob, ((gf + 1) << lg) & 31);
      if ((m0 < 0 || e0 == 0) && (m1 < 0 || e1 == 0)) { skip_rest = true; if (ti != first_t) ok = false; }
    }
    long long wj = kInf;
    int wc = 1 << 20;
    if (ok) { wj = rd_cost(f, r.sse, r.rate + mode_rate(rt, ymode_cost, m, n, DELTA_Y(m))); wc = k * 5 + ti; }
    warp_argmin(wj, wc);
    if (wj < bestJ || (wj == bestJ && wc < bestC)) {
      bestJ = wj; bestC = wc;
      const int wg = (wc % 5) * 2 + wc / 5 - b * G;
      bestEob = __shfl_sync(FULL, r.eob, wg << lg);
    }
  }
  const int best_mode = bestC / 5 == 0 ? m0 : m1;
  const int best_tt = dv_tx_types[bestC % 5];
  long long total = bestJ;
  int uv_best = DC_PRED, eobU = 0, eobV = 0, cfl_js = 0, cfl_idx = 0;
  if (planes > 1) {
    const uint16_t *uvcost = rt->uv_mode[1][best_mode];
    // chroma candidates (oracle): the chosen luma mode with its angle delta, DC_PRED, chroma-from-luma
    int u0 = -1, u1 = -1;
    if ((f->uv_mode_mask >> best_mode) & 1) u0 = best_mode;
    if (best_mode != DC_PRED && (f->uv_mode_mask & 1)) { if (u0 < 0) u0 = DC_PRED; else u1 = DC_PRED; }
    if (u0 < 0) u0 = __ffs(f->uv_mode_mask & 0x1fff) - 1;       // forced single modes (conformance tests)
    long long bestCj = kInf;
    // third candidate: chroma from (source) luma, alpha = clamped integer least-squares fit per plane
    int al_u = 0, al_v = 0, dc_u = 0, dc_v = 0;
    if (f->enable_cfl) {
      cfl_build_ac(sh.src[0], lg, sh.ac);
      dc_u = dc_value_warp(n, sh.above[1] + EOFF, sh.left[1] + EOFF, ha, hl, bd);
      dc_v = dc_value_warp(n, sh.above[2] + EOFF, sh.left[2] + EOFF, ha, hl, bd);
      al_u = cfl_pick_alpha(sh.ac, sh.src[1], N2, dc_u);
      al_v = cfl_pick_alpha(sh.ac, sh.src[2], N2, dc_v);
    }
    const bool cfl_ok = al_u != 0 || al_v != 0;
    const int NC = cfl_ok ? 6 : 4;            // candidate = (k = ci >> 1: rank 0, rank 1, CFL; plane = 1 + (ci & 1))
#pragma unroll 1
    for (int b = 0; b * G < (SYNC ? 6 : NC); b++) {
      if (SYNC) { __syncthreads(); if (b * G >= NC) continue; }
      const int ci = b * G + g;
      const int k = ci >> 1, p = 1 + (ci & 1);
      const int m = k == 0 ? u0 : (k == 1 ? u1 : UV_CFL_PRED);
      const bool ok = ci < NC && m >= 0;
#pragma unroll 1
      for (int cc = b * G; cc < min((b + 1) * G, NC); cc++) {
        const int kk = cc >> 1, pp = 1 + (cc & 1);
        uint16_t *dst = sh.pred + (cc - b * G) * N2;
        if (kk == 2) cfl_predict_block(sh.ac, N2, pp == 1 ? dc_u : dc_v, pp == 1 ? al_u : al_v, bd, dst);
        else {
          const int mm = kk == 0 ? u0 : u1;
          predict_any(sh, ief, mm >= 0 ? mm : 0, DELTA_UV(mm >= 0 ? mm : 0), lg, sh.above[pp], sh.left[pp], ha, hl, bd, dst);
        }
      }
      __syncwarp();
      const TrialOut r = trial_batch<n, HBD>(t, sh, p, sh.src[p], sh.pred + (ok ? g : 0) * N2, chroma_tx_type(ok ? m : 0), best_mode);
      if (ok && (lane & (n - 1)) == 0) { sh.taskD[ci] = r.sse; sh.taskR[ci] = r.rate; sh.taskE[ci] = (uint16_t)r.eob; }
      __syncwarp();
#pragma unroll 1
      for (int kk = 0; kk < NC / 2; kk++) {
        const int cU = kk * 2, cV = kk * 2 + 1;
        if (cV < b * G || cV >= (b + 1) * G) continue;
        const int mm = kk == 0 ? u0 : (kk == 1 ? u1 : UV_CFL_PRED);
        if (mm < 0) continue;
        const uint32_t mr = kk == 2 ? uvcost[UV_CFL_PRED] + cfl_rate(rt, al_u, al_v) : mode_rate(rt, uvcost, mm, n, DELTA_UV(mm));
        const long long J = rd_cost(f, (unsigned long long)sh.taskD[cU] + sh.taskD[cV], sh.taskR[cU] + sh.taskR[cV] + mr);
        if (J < bestCj) { bestCj = J; uv_best = mm; eobU = sh.taskE[cU]; eobV = sh.taskE[cV]; }
      }
      __syncwarp();
    }
    if (uv_best == UV_CFL_PRED) {
      const int su = al_u == 0 ? 0 : (al_u < 0 ? 1 : 2), sv = al_v == 0 ? 0 : (al_v < 0 ? 1 : 2);
      cfl_js = su * 3 + sv - 1;
      cfl_idx = (al_u ? abs(al_u) - 1 : 0) | ((al_v ? abs(al_v) - 1 : 0) << 4);
    }
    total += bestCj;
  }
  if (lane == 0 && out) {
    const int all_zero = bestEob == 0 && (planes == 1 || (eobU == 0 && eobV == 0));
    total += ((long long)f->lambda_q8 * rt->skip[0][all_zero]) >> 8;
    B200BlockDec d;
    d.cost = total; d.y_mode = (uint8_t)best_mode; d.tx_type = (uint8_t)best_tt; d.uv_mode = (uint8_t)uv_best;
    d.all_zero = (uint8_t)all_zero;
    d.angle_y = (uint8_t)(DELTA_Y(best_mode) + 3); d.angle_uv = (uint8_t)(DELTA_UV(uv_best) + 3);
    d.cfl_js = (uint8_t)cfl_js; d.cfl_alpha = (uint8_t)cfl_idx;
    *out = d;
  }
  __syncwarp();
#undef DELTA_Y
#undef DELTA_UV
}

// ---- stage C: closed-loop reconstruction of one decided block by one warp (oracle reconstruct_block) ----------------
template <bool HBD, class SH>
__device__ __noinline__ void reconstruct_block(const TileCtx &t, SH &sh, int x, int y, int lg) {
  const int n = 1 << lg, N2 = n * n, G = 32 >> lg;
  const B200Frame *f = t.f;
  const int lane = lane_id(), g = lane >> lg;
  const int planes = f->planes, bd = f->bit_depth;
  const int mi_r = y >> 2, mi_c = x >> 2, n4 = n >> 2;
  const B200Mi m = t.mi[mi_r * f->mi_cols + mi_c];
#pragma unroll 1
  for (int p = 0; p < planes; p++)
#pragma unroll 1
    for (int q = lane; q < N2; q += 32) sh.src[p][q] = t.src[p][(y + (q >> lg)) * t.stride + x + (q & (n - 1))];
#pragma unroll 1
  for (int i = lane; i < (int)sizeof(sh.lev) / 4; i += 32) reinterpret_cast<uint32_t *>(sh.lev)[i] = 0;
  const int avail = block_avail(t, x, y, n), ha = avail & 1, hl = (avail >> 1) & 1;
#pragma unroll 1
  for (int p = 0; p < planes; p++) build_edges(t, sh, t.rec[p], p, x, y, n, avail);
  __syncwarp();
  const long long coff = b200_coef_offset(f, mi_r, mi_c);
  int any = 0;
  // CFL chroma needs the reconstructed luma of this block: luma goes alone first, then both chroma planes
  const bool cfl = planes > 1 && m.uv_mode == UV_CFL_PRED;
  int au = 0, av = 0;
  if (cfl) {
    const int js = (m.flags >> 1) & 7, su = (js + 1) / 3, sv = (js + 1) % 3;
    au = su == 0 ? 0 : (su == 1 ? -1 : 1) * ((m.cfl & 15) + 1);
    av = sv == 0 ? 0 : (sv == 1 ? -1 : 1) * ((m.cfl >> 4) + 1);
  }
#pragma unroll 1
  for (int p0 = 0; p0 < planes;) {              // plane p = p0 + g handled by lane group g
    const int cnt = (cfl && p0 == 0) ? 1 : min(G, planes - p0);
#pragma unroll 1
    for (int pp = p0; pp < p0 + cnt; pp++) {
      uint16_t *dst = sh.pred + (pp - p0) * N2;
      if (cfl && pp > 0)
        cfl_predict_block(sh.ac, N2, dc_value_warp(n, sh.above[pp] + EOFF, sh.left[pp] + EOFF, ha, hl, bd), pp == 1 ? au : av, bd, dst);
      else
      {
        int ief = -1;
        if (f->enable_intra_edge_filter) {     // get_filter_type (spec 7.11.2.8): a smooth mode above or to the left
          ief = 0;
          if (ha) { const B200Mi nb = t.mi[(mi_r - 1) * f->mi_cols + mi_c]; const int nm = pp == 0 ? nb.y_mode : nb.uv_mode; if (nm >= SMOOTH_PRED && nm <= SMOOTH_H_PRED) ief = 1; }
          if (hl) { const B200Mi nb = t.mi[mi_r * f->mi_cols + mi_c - 1]; const int nm = pp == 0 ? nb.y_mode : nb.uv_mode; if (nm >= SMOOTH_PRED && nm <= SMOOTH_H_PRED) ief = 1; }
        }
        predict_any(sh, ief, pp == 0 ? m.y_mode : m.uv_mode, (pp == 0 ? (m.angle & 15) : (m.angle >> 4)) - 3, lg, sh.above[pp], sh.left[pp], ha, hl, bd, dst);
      }
    }
    __syncwarp();
    const int p = p0 + g;
    const bool ok = g < cnt;
    const int pe = ok ? p : 0;
    const TrialOut r = trial_dispatch<HBD>(n, t, sh, pe, sh.src[pe], sh.pred + (ok ? g : 0) * N2,
                                           pe == 0 ? m.tx_type : chroma_tx_type(m.uv_mode), m.y_mode);
    if (ok) {
      // every lane of the group writes its share of the block: reconstruction and scan-ordered levels
      const int l = lane & (n - 1);
#pragma unroll 1
      for (int q = l; q < N2; q += n) {
        t.rec[p][(y + (q >> lg)) * t.stride + x + (q & (n - 1))] = sh.rec[g * N2 + q];
        t.coef[p][coff + q] = sh.qc[g * N2 + q];
      }
      if (l == 0) t.eob[p][mi_r * f->mi_cols + mi_c] = (uint16_t)r.eob;
    }
    any |= __ballot_sync(FULL, ok && r.eob != 0) != 0;
    __syncwarp();
    if (cfl && p0 == 0) cfl_build_ac(sh.rec, lg, sh.ac);
    p0 += cnt;
  }
  B200Mi mo = m;
  if (!any) mo.flags |= 1;
  for (int i = lane; i < n4 * n4; i += 32) t.mi[(mi_r + (i >> (lg - 2))) * f->mi_cols + mi_c + (i & (n4 - 1))] = mo;
  __syncwarp();
}

}  // namespace

// ---- stage A.1: SATD of the 13 intra modes for every block of level lv, predicted from SOURCE neighbours.
// One thread per (block, plane group): group 0 = luma, group 1 = U + V.  out: [block][2][13] uint32.
// One item = one block of level lv (luma SATD of the 13 modes + angle refinement).  A CTA owns kSatdItems consecutive items: it first compacts the items
// that survive the top-down pruning into shared memory, then its 128 threads walk the compacted list, so warps stay
// full even when most of a level is pruned.
constexpr int kSatdItems = 512;
__device__ __forceinline__ void satd_item(const B200RdoJob &job, int lv, int gid, int *sh_d);

__global__ void __launch_bounds__(128) b200_satd_kernel(const B200RdoJob *jobs, int lv, int items /* per CTA: 128 .. kSatdItems */) {
  __shared__ int sh_d[128 * 17];     // per-thread 16 residuals (+1 pad: conflict-free), written by a rolled loop
  __shared__ uint16_t s_list[kSatdItems];
  __shared__ int s_n;
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
  const int n = 4 << lv;
  if (n < min(f->partition_min, 8) || n > f->partition_max) return;   // 8x8 always: forced splits at the frame edge
  const int cols = b200_level_cols(f, lv), rows = b200_level_rows(f, lv);
  const int total = cols * rows;
  const int base = blockIdx.x * items;
  if (base >= total) return;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
#pragma unroll 1
  for (int k = 0; k < items / 128; k++) {
    const int it = k * 128 + threadIdx.x, gid = base + it;
    bool active = gid < total;
    if (active && lv < B200_LEVELS - 1) {
      const int blk = gid % (cols * rows);
      active = !b200_pruned_by_parent(f, lv, blk % cols, blk / cols, job.dec[lv + 1]);   // top-down pruning
    }
    const unsigned m = __ballot_sync(FULL, active);
    int off = 0;
    if (lane_id() == 0 && m) off = atomicAdd(&s_n, __popc(m));
    off = __shfl_sync(FULL, off, 0);
    if (active) s_list[off + __popc(m & ((1u << lane_id()) - 1))] = (uint16_t)it;
  }
  __syncthreads();
  const int cnt = s_n;
#pragma unroll 1
  for (int i = threadIdx.x; i < cnt; i += 128) satd_item(job, lv, base + s_list[i], sh_d);
}

// never-pruned top level: one item per thread, no compaction
__global__ void __launch_bounds__(128) b200_satd_direct_kernel(const B200RdoJob *jobs, int lv) {
  __shared__ int sh_d[128 * 17];
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
  const int n = 4 << lv;
  if (n < min(f->partition_min, 8) || n > f->partition_max) return;
  const int total = b200_level_cols(f, lv) * b200_level_rows(f, lv);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  if (lv < B200_LEVELS - 1) {
    const int cols = b200_level_cols(f, lv), blk = gid % (cols * b200_level_rows(f, lv));
    if (b200_pruned_by_parent(f, lv, blk % cols, blk / cols, job.dec[lv + 1])) return;
  }
  satd_item(job, lv, gid, sh_d);
}

__device__ __forceinline__ void satd_item(const B200RdoJob &job, int lv, int gid, int *sh_d) {
  const B200Frame *f = &job.frame;
  const int n = 4 << lv, lg = 2 + lv;
  const int cols = b200_level_cols(f, lv);
  constexpr int grp = 0;            // luma only: the chroma candidates are derived from the luma decision
  const int blk = gid;
  const int bx = blk % cols, by = blk / cols, x = bx * n, y = by * n;
  TileCtx t;
  ctx_from_job(job, t);
  tile_of_unit(f, y >> 2, x >> 2, t);
  const int st = t.stride, bd = f->bit_depth;
  const int mi_r = y >> 2, mi_c = x >> 2, n4 = n >> 2;
  const int ha = mi_r > t.tr0, hl = mi_c > t.tc0;
  const int max_x = f->mi_cols * 4 - 1, max_y = f->mi_rows * 4 - 1;
  const int base = 1 << (bd - 1);
  const int have_ar = ha && b200_unit_before(mi_r - 1, mi_c + n4, mi_r, mi_c, t.tr0, t.tr1, t.tc0, t.tc1);
  const int have_bl = hl && b200_unit_before(mi_r + n4, mi_c - 1, mi_r, mi_c, t.tr0, t.tr1, t.tc0, t.tc1);
  const int mask = grp == 0 ? f->y_mode_mask : f->uv_mode_mask;
  // acc[0..12]: SATD of the 13 nominal modes; acc[13..18]: the refined directional mode at deltas -3,-2,-1,1,2,3.
  // Both passes run through ONE rolled copy of the edge / predictor / Hadamard code (instruction-cache footprint).
  uint32_t acc[19];
#pragma unroll
  for (int m = 0; m < 19; m++) acc[m] = 0;
  uint16_t Ab[EOFF + 66], Lb[EOFF + 66];       // n <= 32
  uint16_t *A = Ab + EOFF, *L = Lb + EOFF;
  const int p_end = grp == 0 ? 1 : 3;
  int md = -1;
  uint32_t smd = 0;
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      // angle refinement of the best directional mode (oracle refine_angle)
      if (!(f->fine_directional_intra && n >= 8)) break;
#pragma unroll 1
      for (int m = V_PRED; m <= D67_PRED; m++)
        if (((mask >> m) & 1) && (md < 0 || acc[m] < smd)) { md = m; smd = acc[m]; }
      if (md < 0) break;
    }
    const int ncand = pass ? 6 : 13;
#pragma unroll 1
    for (int p = grp; p < p_end; p++) {
      const uint16_t *r = job.src[p];
      if (pass == 0 || p_end - grp > 1) {     // single-plane group: the edges of pass 0 are still valid in pass 1
        // source-neighbour edges (same rules as build_edges, one thread)
#pragma unroll 1
        for (int i = 0; i < 2 * n; i++) {
          uint16_t a, l;
          if (!ha && hl) a = r[y * st + x - 1];
          else if (!ha && !hl) a = (uint16_t)(base - 1);
          else { int xx = (i < n || have_ar) ? x + i : x + n - 1; if (xx > max_x) xx = max_x; a = r[(y - 1) * st + xx]; }
          if (!hl && ha) l = r[(y - 1) * st + x];
          else if (!hl && !ha) l = (uint16_t)(base + 1);
          else { int yy = (i < n || have_bl) ? y + i : y + n - 1; if (yy > max_y) yy = max_y; l = r[yy * st + x - 1]; }
          A[i] = a; L[i] = l;
        }
        uint16_t tl;
        if (ha && hl) tl = r[(y - 1) * st + x - 1];
        else if (ha) tl = r[(y - 1) * st + x];
        else if (hl) tl = r[y * st + x - 1];
        else tl = (uint16_t)base;
        A[-1] = tl; L[-1] = tl;
      }
      const int dc = dc_value(n, A, L, ha, hl, bd);
#pragma unroll 1
      for (int sb = 0; sb < (n4 * n4); sb++) {
        const int sy = (sb >> (lg - 2)) * 4, sx = (sb & (n4 - 1)) * 4;
        int s[16];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) s[i * 4 + j] = r[(y + sy + i) * st + x + sx + j];
#pragma unroll 1
        for (int c = 0; c < ncand; c++) {
          if (pass == 0 && !((mask >> c) & 1)) continue;
          const int m = pass ? md : c, delta = pass ? (c < 3 ? c - 3 : c - 2) : 0;
          int tt[16], d[16];
#pragma unroll 1
          for (int k = 0; k < 16; k++) sh_d[threadIdx.x * 17 + k] = pred_px(m, delta, n, sy + (k >> 2), sx + (k & 3), A, L, dc);
#pragma unroll
          for (int k = 0; k < 16; k++) d[k] = s[k] - sh_d[threadIdx.x * 17 + k];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int a = d[i * 4] + d[i * 4 + 1], b = d[i * 4] - d[i * 4 + 1], c2 = d[i * 4 + 2] + d[i * 4 + 3], e = d[i * 4 + 2] - d[i * 4 + 3];
            tt[i * 4] = a + c2; tt[i * 4 + 1] = b + e; tt[i * 4 + 2] = a - c2; tt[i * 4 + 3] = b - e;
          }
          uint32_t total = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int a = tt[j] + tt[4 + j], b = tt[j] - tt[4 + j], c2 = tt[8 + j] + tt[12 + j], e = tt[8 + j] - tt[12 + j];
            total += (uint32_t)(abs(a + c2) + abs(b + e) + abs(a - c2) + abs(b - e));
          }
          acc[pass * 13 + c] += total;
        }
      }
    }
  }
  uint32_t word = 255u | (3u << 8);
  if (md >= 0) {
    const uint16_t *drate = job.rate->angle_delta[md - V_PRED];
    long long best = (long long)smd * 64 + (((long long)f->lambda_satd_q8 * drate[3]) >> 8);
    int best_d = 0;
    uint32_t best_s = smd;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
      const int d = k < 3 ? k - 3 : k - 2;
      const long long c = (long long)acc[13 + k] * 64 + (((long long)f->lambda_satd_q8 * drate[d + 3]) >> 8);
      if (c < best) { best = c; best_d = d; best_s = acc[13 + k]; }
    }
    acc[md] = best_s;
    word = (uint32_t)md | ((uint32_t)(best_d + 3) << 8);
  }
  uint32_t *out = job.satd[lv] + (size_t)blk * B200_SATD_WORDS;
#pragma unroll 1
  for (int m = 0; m < 13; m++) out[m] = acc[m];
  out[13] = word;
}

// ---- stage A.2: full trials.  One single-warp CTA loops over blocks of level lv (grid-stride), grid.y = job.
// W warps per CTA.  W == 1: free-running single-warp CTAs (grid-stride over the blocks).  W > 1 (the never-pruned top
// level): the W warps of a CTA analyse W different blocks in lock-step (analyze_block<.., SYNC>).
constexpr int kAnalyzeWarpsTop = 4;      // 16x16 as the top level
constexpr int kAnalyzeWarpsTop32 = 2;    // 32x32 as the top level (22.9 KB of shared memory per warp)
template <int LV, bool HBD, int W>
__global__ void __launch_bounds__(kThreads * W, LV == 3 ? 4 : (LV == 2 ? 16 : 24) / W) b200_analyze_kernel(const B200RdoJob *jobs) {
  typedef SharedT<(LV == 3 ? 32 : 16)> SH;
  __shared__ SH shs[W];
  SH &sh = shs[threadIdx.x >> 5];
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
  constexpr int lv = LV, n = 4 << LV;
  if (n < min(f->partition_min, 8) || n > f->partition_max) return;   // 8x8 always: forced splits at the frame edge
  TileCtx t;
  ctx_from_job(job, t);
  const int cols = b200_level_cols(f, lv), total = cols * b200_level_rows(f, lv);
  if (W > 1) {
    const int per_iter = gridDim.x * W, iters = (total + per_iter - 1) / per_iter;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
      const int want = (it * gridDim.x + blockIdx.x) * W + (threadIdx.x >> 5);
      const int blk = want < total ? want : total - 1;             // padding warps repeat the last block, output discarded
      const int x = (blk % cols) * n, y = (blk / cols) * n;
      {   // the next block of this warp: pull its source rows into L2 while this one is analysed
        const int nxt = want + per_iter;
        if (nxt < total) {
          const int nx = (nxt % cols) * n, ny = (nxt / cols) * n;
          for (int i = lane_id(); i < n * f->planes; i += 32) B200_PREFETCH_L2(t.src[i / n] + (size_t)(ny + i % n) * t.stride + nx);
        }
      }
      tile_of_unit(f, y >> 2, x >> 2, t);
      analyze_block<2 + LV, HBD, true>(t, sh, x, y, job.satd[lv] + (size_t)blk * B200_SATD_WORDS, want < total ? job.dec[lv] + blk : nullptr);
    }
    return;
  }
#pragma unroll 1
  for (int blk = blockIdx.x; blk < total; blk += gridDim.x) {
    const int x = (blk % cols) * n, y = (blk / cols) * n;
    if (LV < B200_LEVELS - 1 && b200_pruned_by_parent(f, lv, blk % cols, blk / cols, job.dec[LV < B200_LEVELS - 1 ? LV + 1 : LV])) {
      if (lane_id() == 0) {   // top-down pruning: never split down to this block
        B200BlockDec d;
        d.cost = B200_COST_PRUNED; d.y_mode = 0; d.tx_type = 0; d.uv_mode = 0; d.all_zero = 0; d.angle_y = 0; d.angle_uv = 0;
        d.cfl_js = 0; d.cfl_alpha = 0;
        job.dec[lv][blk] = d;
      }
      continue;
    }
    tile_of_unit(f, y >> 2, x >> 2, t);
    analyze_block<2 + LV, HBD, false>(t, sh, x, y, job.satd[lv] + (size_t)blk * B200_SATD_WORDS, job.dec[lv] + blk);
  }
}

// ---- stage B: partition decision, one thread per 16x16 region (oracle orc_decide_region)
namespace {
__device__ __forceinline__ long long part_cost0(const B200Frame *f, const B200RateTables *rt, int mi_r, int mi_c, int n, int split) {
  const int bsl = ilog2_dev(n) - 2, hbs = n >> 3;
  if (!((mi_r + hbs) < f->mi_rows && (mi_c + hbs) < f->mi_cols)) return 0;
  return ((long long)f->lambda_q8 * rt->partition[(bsl - 1) * 4][split ? 3 : 0]) >> 8;
}
__device__ __forceinline__ void set_mi(const B200RdoJob &job, int x, int y, int n, const B200BlockDec &d) {
  const B200Frame *f = &job.frame;
  B200Mi m;
  m.bsize = (uint8_t)(n == 4 ? BLOCK_4X4 : n == 8 ? BLOCK_8X8 : n == 16 ? BLOCK_16X16 : BLOCK_32X32);
  m.y_mode = d.y_mode; m.uv_mode = d.uv_mode;
  m.angle = (uint8_t)((d.angle_y ? d.angle_y : 3) | ((d.angle_uv ? d.angle_uv : 3) << 4));
  m.tx_size = (uint8_t)txs_of(n); m.tx_type = d.tx_type; m.flags = 0; m.cfl = 0;
  if (d.uv_mode == UV_CFL_PRED) { m.flags = (uint8_t)(d.cfl_js << 1); m.cfl = d.cfl_alpha; }
  for (int i = 0; i < n / 4; i++)
    for (int j = 0; j < n / 4; j++) job.mi[((y >> 2) + i) * f->mi_cols + (x >> 2) + j] = m;
}
__device__ __forceinline__ const B200BlockDec &dec_at(const B200RdoJob &job, int lv, int x, int y) {
  const int n = 4 << lv;
  return job.dec[lv][(y / n) * b200_level_cols(&job.frame, lv) + (x / n)];
}
}  // namespace

// oracle decide_block(): best RD cost of block (LV, x, y), NONE against SPLIT recursively; commit writes the decisions
template <int LV>
__device__ long long decide_block(const B200RdoJob &job, int x, int y, bool commit) {
  const B200Frame *f = &job.frame;
  const B200RateTables *rt = job.rate;
  const int W = f->mi_cols * 4, H = f->mi_rows * 4;
  constexpr int n = 4 << LV;
  if (x >= W || y >= H) return 0;
  const bool inside = (x + n <= W) && (y + n <= H);
  const bool can_none = inside && n <= f->partition_max;
  const bool can_split = LV > 0 && (f->partition_min < n || !inside);
  long long j_none = 0, j_split = 0;
  if (can_none) j_none = dec_at(job, LV, x, y).cost + (LV > 0 ? part_cost0(f, rt, y >> 2, x >> 2, n, 0) : 0);
  if (LV > 0 && can_split) {
    j_split = part_cost0(f, rt, y >> 2, x >> 2, n, 1);
#pragma unroll 1
    for (int k = 0; k < 4; k++) j_split += decide_block<(LV > 0 ? LV - 1 : 0)>(job, x + (k & 1) * (n >> 1), y + (k >> 1) * (n >> 1), false);
  }
  const bool none = can_none && (!can_split || j_none <= j_split);
  if (commit) {
    if (none) set_mi(job, x, y, n, dec_at(job, LV, x, y));
    else if (LV > 0) {
#pragma unroll 1
      for (int k = 0; k < 4; k++) decide_block<(LV > 0 ? LV - 1 : 0)>(job, x + (k & 1) * (n >> 1), y + (k >> 1) * (n >> 1), true);
    }
  }
  return none ? j_none : j_split;
}

// one thread per root region: 32x32 when the partition range reaches it, else 16x16
__global__ void __launch_bounds__(128) b200_decide_kernel(const B200RdoJob *jobs) {
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
  const int W = f->mi_cols * 4, H = f->mi_rows * 4;
  const int root = f->partition_max >= 32 ? 3 : 2, rn = 4 << root;
  const int rcols = (W + rn - 1) / rn, rrows = (H + rn - 1) / rn;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rcols * rrows) return;
  const int x = (gid % rcols) * rn, y = (gid / rcols) * rn;
  if (root == 3) decide_block<3>(job, x, y, true);
  else decide_block<2>(job, x, y, true);
}

// ---- stage C: one CTA per tile, kReconWarps warps on a superblock-row wavefront.  Warp w owns the tile's SB rows
// w, w + W, ...; it walks a row left to right (the decided blocks of each SB in z-order) and may start SB (R, C) once
// row R - 1 has finished SB C + 1: the above-right superblock is the furthest reconstruction any intra edge of this
// SB can read (its bottom-left 64x64 neighbour is never decoded before it).  Progress counters live in shared memory.
constexpr int kReconWarps = 4;
constexpr int kMaxSbRows = 1024;     // 65536 / 64
#ifdef B200_EMU
#define B200_SPIN_WAIT() emu::yield()
#else
#define B200_SPIN_WAIT() __nanosleep(40)
#endif
// The chain stalls on the first touch of every block's source samples (profiles/: 27 % of the active samples of the
// reconstruction kernel): pull the source rows of superblock (mi row r, mi col c) into L2 ahead of time, one 128-byte
// line per lane and iteration (64 samples of 2 bytes = one line per row and plane).
__device__ __forceinline__ void prefetch_sb_source(const TileCtx &t, int r, int c) {
  const B200Frame *f = t.f;
  if (r >= t.tr1 || c >= t.tc1) return;
  const int rows = min(64, (f->mi_rows - r) * 4);
#pragma unroll 1
  for (int i = lane_id(); i < rows * f->planes; i += 32) {
    const int p = i / rows, y = r * 4 + i % rows;
    B200_PREFETCH_L2(t.src[p] + (size_t)y * t.stride + c * 4);
  }
}
// NMAX = 32 (partition range up to 32x32): 2 warps per tile, 22.9 KB of shared memory each
template <bool HBD, int NMAX>
__global__ void __launch_bounds__(kThreads * (NMAX == 32 ? 2 : kReconWarps), NMAX == 32 ? 4 : 4) b200_recon_kernel(const B200RdoJob *jobs, const B200TileRef *tiles) {
  constexpr int kWarps = NMAX == 32 ? 2 : kReconWarps;
  typedef SharedT<NMAX> SH;
  __shared__ SH shs[kWarps];
  // superblocks finished in each SB row of this tile: 16-bit counters (<= 1024 columns), two per word, touched only by
  // shared-memory atomics (release: fence + atomicAdd by the producer warp, acquire: atomic read + fence by the consumer)
  __shared__ unsigned int sb_done[kMaxSbRows / 2];
  const B200TileRef tr = tiles[blockIdx.x];
  const B200RdoJob &job = jobs[tr.job];
  TileCtx t;
  ctx_from_job(job, t);
  const B200Frame *f = t.f;
  t.tr0 = f->tile_row_start_sb[tr.tile_row] * 16;
  t.tr1 = min(f->tile_row_start_sb[tr.tile_row + 1] * 16, f->mi_rows);
  t.tc0 = f->tile_col_start_sb[tr.tile_col] * 16;
  t.tc1 = min(f->tile_col_start_sb[tr.tile_col + 1] * 16, f->mi_cols);
  const int sb_rows = (t.tr1 - t.tr0 + 15) >> 4, sb_cols = (t.tc1 - t.tc0 + 15) >> 4;
  for (int i = threadIdx.x; i < (sb_rows + 1) / 2; i += blockDim.x) sb_done[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  SH &sh = shs[warp];
  const int n_warps = blockDim.x >> 5;       // <= kWarps (1 when the wavefront is switched off for debugging)
#pragma unroll 1
  for (int R = warp; R < sb_rows; R += n_warps) {
    const int r = t.tr0 + R * 16;
    prefetch_sb_source(t, r, t.tc0);
#pragma unroll 1
    for (int C = 0; C < sb_cols; C++) {
      prefetch_sb_source(t, r, t.tc0 + (C + 1) * 16);        // the next superblock of this row, while this one is coded
      if (R > 0) {
        const int need = min(C + 2, sb_cols);
        while ((int)((atomicAdd(&sb_done[(R - 1) >> 1], 0u) >> (16 * ((R - 1) & 1))) & 0xffffu) < need) B200_SPIN_WAIT();
        __syncwarp();
        __threadfence_block();
      }
      const int c = t.tc0 + C * 16;
#pragma unroll 1
      for (int z = 0; z < 256;) {          // 4x4 units of the superblock in z-order; skip over whole blocks
        int ur = 0, uc = 0;
        for (int b = 0; b < 4; b++) { uc |= ((z >> (2 * b)) & 1) << b; ur |= ((z >> (2 * b + 1)) & 1) << b; }
        const int mr = r + ur, mc = c + uc;
        if (mr >= f->mi_rows || mc >= f->mi_cols) { z++; continue; }
        const int bs = t.mi[mr * f->mi_cols + mc].bsize;
        const int lg = bs == BLOCK_4X4 ? 2 : (bs == BLOCK_8X8 ? 3 : (bs == BLOCK_16X16 ? 4 : 5));
        reconstruct_block<HBD>(t, sh, mc * 4, mr * 4, lg);
        z += 1 << (2 * (lg - 2));           // a block covers (n/4)^2 consecutive z indices
      }
      __threadfence_block();               // the reconstruction of this SB is visible before the counter moves
      __syncwarp();
      if (lane_id() == 0) atomicAdd(&sb_done[R >> 1], 1u << (16 * (R & 1)));
    }
  }
}

#ifndef B200_EMU
// Stages A + B go to `stream`; stage C goes to `stream_recon` (may be the same stream) after `analysis_done` (an event
// recorded behind stage B, may be null when both streams are the same).  The caller overlaps the low-parallelism
// reconstruction chain of one chunk with the analysis of the next by giving stream_recon a higher priority.
extern "C" int b200_rdo_launch(const B200RdoJob *d_jobs, int n_jobs, const B200TileRef *d_tiles, int n_tiles, int max_blocks4,
                               int max_regions, int high_bitdepth, int top_level, void *stream, void *stream_recon, void *analysis_done, int *launched) {
  if (n_tiles <= 0 || n_jobs <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int n_launch = 0;
  // function attributes are per device: set them once per device, under a lock (concurrent callers, several GPUs)
  static std::mutex cfg_mu;
  static bool cfg_done[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  std::lock_guard<std::mutex> cfg_lock(cfg_mu);
  bool &configured = cfg_done[cur_dev & 63];
  if (!configured) {
#define B200_CARVE(k) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
    B200_CARVE((b200_analyze_kernel<0, false, 1>)); B200_CARVE((b200_analyze_kernel<1, false, 1>)); B200_CARVE((b200_analyze_kernel<2, false, 1>));
    B200_CARVE((b200_analyze_kernel<0, true, 1>)); B200_CARVE((b200_analyze_kernel<1, true, 1>)); B200_CARVE((b200_analyze_kernel<2, true, 1>));
    B200_CARVE((b200_analyze_kernel<2, false, kAnalyzeWarpsTop>)); B200_CARVE((b200_analyze_kernel<2, true, kAnalyzeWarpsTop>));
    B200_CARVE((b200_analyze_kernel<3, false, kAnalyzeWarpsTop32>)); B200_CARVE((b200_analyze_kernel<3, true, kAnalyzeWarpsTop32>));
    B200_CARVE((b200_recon_kernel<false, 16>)); B200_CARVE((b200_recon_kernel<true, 16>));
    B200_CARVE((b200_recon_kernel<false, 32>)); B200_CARVE((b200_recon_kernel<true, 32>));
    configured = true;
  }
  const int top = top_level >= 3 ? 3 : 2;
  for (int lv = top; lv >= 0; lv--) {   // large to small: every level is pruned by its parent level
    const int nblk = max_blocks4 >> (2 * lv);
    if (nblk == 0) continue;
    // the top level is never pruned: one item per thread; lower levels are mostly pruned: compact 512 items per CTA
    if (lv == top) {
      dim3 gs((nblk + 127) / 128, n_jobs);
      b200_satd_direct_kernel<<<gs, 128, 0, st>>>(d_jobs, lv);
      n_launch++;
    } else {
      dim3 gs((nblk + kSatdItems - 1) / kSatdItems, n_jobs);
      b200_satd_kernel<<<gs, 128, 0, st>>>(d_jobs, lv, kSatdItems);
      n_launch++;
    }
    // enough warps to fill the machine several times over, grid-stride inside
    const int per_job = min(nblk, max(1, (148 * 16 * 4) / n_jobs));
    dim3 ga(per_job, n_jobs);
#define B200_ANALYZE(LV, W, GRID) \
    do { if (high_bitdepth) b200_analyze_kernel<LV, true, W><<<GRID, kThreads * W, 0, st>>>(d_jobs); \
         else b200_analyze_kernel<LV, false, W><<<GRID, kThreads * W, 0, st>>>(d_jobs); } while (0)
    if (lv == 3) {
      dim3 gt((per_job + kAnalyzeWarpsTop32 - 1) / kAnalyzeWarpsTop32, n_jobs);
      B200_ANALYZE(3, kAnalyzeWarpsTop32, gt);
    } else if (lv == 2 && top == 2) {      // never pruned: lock-step CTAs
      dim3 gt((per_job + kAnalyzeWarpsTop - 1) / kAnalyzeWarpsTop, n_jobs);
      B200_ANALYZE(2, kAnalyzeWarpsTop, gt);
    } else if (lv == 2) B200_ANALYZE(2, 1, ga);
    else if (lv == 1) B200_ANALYZE(1, 1, ga);
    else B200_ANALYZE(0, 1, ga);
    n_launch++;
  }
  // one thread per ROOT REGION: ceil(W / rn) * ceil(H / rn) per job (larger than n_mi / 16 when the coded size is
  // 8 mod 16), computed by the caller over all jobs
  dim3 gd((max_regions + 127) / 128, n_jobs);
  b200_decide_kernel<<<gd, 128, 0, st>>>(d_jobs);
  n_launch += 2;      // partition decision + the reconstruction chain below
  cudaStream_t sr = stream_recon ? (cudaStream_t)stream_recon : st;
  if (analysis_done) {
    cudaEventRecord((cudaEvent_t)analysis_done, st);
    if (sr != st) cudaStreamWaitEvent(sr, (cudaEvent_t)analysis_done, 0);
  }
  static const bool serial_chain = getenv("B200AVIF_DEBUG_SERIAL_RECON") != nullptr;      // debugging aid: one warp per tile
  if (top == 3) {
    const int bt = kThreads * (serial_chain ? 1 : 2);
    if (high_bitdepth) b200_recon_kernel<true, 32><<<n_tiles, bt, 0, sr>>>(d_jobs, d_tiles);
    else b200_recon_kernel<false, 32><<<n_tiles, bt, 0, sr>>>(d_jobs, d_tiles);
  } else {
    const int bt = kThreads * (serial_chain ? 1 : kReconWarps);
    if (high_bitdepth) b200_recon_kernel<true, 16><<<n_tiles, bt, 0, sr>>>(d_jobs, d_tiles);
    else b200_recon_kernel<false, 16><<<n_tiles, bt, 0, sr>>>(d_jobs, d_tiles);
  }
  if (launched) *launched = n_launch;
  return (int)cudaGetLastError();
}
#else
static const B200RdoJob *g_emu_jobs;
static const B200TileRef *g_emu_tiles;
static int g_emu_lv, g_emu_top;
static void emu_satd(void *) {
  if (g_emu_lv == g_emu_top) b200_satd_direct_kernel(g_emu_jobs, g_emu_lv);
  else b200_satd_kernel(g_emu_jobs, g_emu_lv, kSatdItems);
}
static void emu_analyze(void *) {
  const bool hbd = g_emu_jobs[0].frame.bit_depth > 8;
#define EMU_ANALYZE(LV, W) do { if (hbd) b200_analyze_kernel<LV, true, W>(g_emu_jobs); else b200_analyze_kernel<LV, false, W>(g_emu_jobs); } while (0)
  if (g_emu_lv == 0) EMU_ANALYZE(0, 1);
  else if (g_emu_lv == 1) EMU_ANALYZE(1, 1);
  else if (g_emu_lv == 2 && g_emu_top == 2) EMU_ANALYZE(2, kAnalyzeWarpsTop);
  else if (g_emu_lv == 2) EMU_ANALYZE(2, 1);
  else EMU_ANALYZE(3, kAnalyzeWarpsTop32);
}
static void emu_decide(void *) { b200_decide_kernel(g_emu_jobs); }
static void emu_recon(void *) {
  const bool hbd = g_emu_jobs[0].frame.bit_depth > 8;
  if (g_emu_top == 3) { if (hbd) b200_recon_kernel<true, 32>(g_emu_jobs, g_emu_tiles); else b200_recon_kernel<false, 32>(g_emu_jobs, g_emu_tiles); }
  else { if (hbd) b200_recon_kernel<true, 16>(g_emu_jobs, g_emu_tiles); else b200_recon_kernel<false, 16>(g_emu_jobs, g_emu_tiles); }
}
extern "C" int b200_rdo_launch_emu(const B200RdoJob *jobs, const B200TileRef *tiles, int n_tiles) {
  g_emu_jobs = jobs; g_emu_tiles = tiles;
  const B200Frame *f = &jobs[0].frame;
  g_emu_top = f->partition_max >= 32 ? 3 : 2;
  for (g_emu_lv = g_emu_top; g_emu_lv >= 0; g_emu_lv--) {
    const int nblk = (f->mi_rows >> g_emu_lv) * (f->mi_cols >> g_emu_lv);
    if (nblk == 0) continue;
    const int items = g_emu_lv == g_emu_top ? 128 : kSatdItems;
    emu::launch(emu_satd, nullptr, (unsigned)((nblk + items - 1) / items), 128);
    const int w = g_emu_lv == 3 ? kAnalyzeWarpsTop32 : ((g_emu_lv == 2 && g_emu_top == 2) ? kAnalyzeWarpsTop : 1);
    emu::launch(emu_analyze, nullptr, (unsigned)std::min(nblk, w > 1 ? 3 : 7), kThreads * w);
  }
  const int rn4 = g_emu_top == 3 ? 8 : 4;
  const int regions = ((f->mi_cols + rn4 - 1) / rn4) * ((f->mi_rows + rn4 - 1) / rn4);
  emu::launch(emu_decide, nullptr, (unsigned)((regions + 127) / 128), 128);
  emu::launch(emu_recon, nullptr, (unsigned)n_tiles, kThreads * (g_emu_top == 3 ? 2 : kReconWarps));
  return 0;
}
#endif
