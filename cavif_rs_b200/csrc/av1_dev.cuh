// Device code shared by the RDO kernels (rdo.cu: open-loop analysis, recon.cu: closed-loop reconstruction chain):
// AV1 tables, intra predictors (spec 7.11.2) incl. the intra edge filter / upsampling, chroma from luma, the 1-D
// transform kernels and the lane-batched trial (forward transform -> quantise -> rate -> dequantise -> inverse
// transform -> reconstruction -> SSE).  Everything that touches samples is generic over the sample type:
// uint8_t planes for 8-bit video, uint16_t for 10-bit (native sample width: half the HBM / L2 / shared-memory
// footprint for 8-bit video).  Bit-exact CPU restatement: oracle/av1_encode.c.
#ifndef B200_AV1_DEV_CUH_
#define B200_AV1_DEV_CUH_
#ifdef B200_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdlib.h>

#include "records.h"
#include "rdo.h"
#ifdef B200_EMU
#define B200_MARK(name, st) ((void)0)
#else
#define B200_MARK(name, st) b200_mark(name, (void *)(st))
#endif

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
template <bool HBD> struct PixOf { typedef uint16_t type; };
template <> struct PixOf<false> { typedef uint8_t type; };

// inverse scan tables (4x4 | 8x8 | 16x16) in GLOBAL memory: indexed per lane (divergent), which the constant
// cache would serialise; __ldg keeps them L1 resident.
__device__ __align__(16) const uint8_t dv_iscan_g[336] = {
#include "iscan_init.inc"
};
__device__ __align__(16) const uint16_t dv_iscan32_g[1024] = {
#include "iscan32_init.inc"
};
__constant__ const uint8_t dv_sm_weights[4 + 8 + 16 + 32] = {255, 149, 85, 64, 255, 197, 146, 105, 73, 50, 37, 32,
                                                             255, 225, 196, 170, 145, 123, 102, 84, 68, 54, 43, 33, 26, 20, 17, 16,
                                                             255, 240, 225, 210, 196, 182, 169, 157, 145, 133, 122, 111, 101, 92, 83, 74,
                                                             66, 59, 52, 45, 39, 34, 29, 25, 21, 17, 14, 12, 10, 9, 8, 8};
__constant__ const uint8_t dv_tx_types[B200_TX_TYPES] = {B200_TX_TYPE_LIST};

// Per-warp working set.  NMAX = largest block size the kernel instance handles; PIX = sample type of the planes
// (uint8_t for 8-bit video, uint16_t for 10-bit): 16 / u8: 6.6 KB, 16 / u16: 8.7 KB, 32 / u8: 16.1 KB, 32 / u16: 22.9 KB.
template <int NMAX, class PIX>
struct SharedT {
  typedef PIX Pix;
  static constexpr int kMaxN = NMAX;
  static constexpr int kPredStride = NMAX * NMAX;          // one prediction slot (luma: two ranked modes, two slots)
  alignas(16) PIX src[3][NMAX * NMAX];
  PIX above[3][EOFF + 2 * NMAX + 2], left[3][EOFF + 2 * NMAX + 2];
  PIX ef[2][EOFF + 2 * NMAX + 8];   // intra edge filter: filtered / upsampled private copy of one plane's edges
  PIX pred[2 * NMAX * NMAX];   // luma: predictions of the two ranked modes; chroma: up to G (mode rank, plane) predictions
  int32_t buf[32 * (NMAX + 1)];     // G groups x N rows x (N+1): transform transpose staging
  alignas(16) uint8_t lev[NMAX == 32 ? 1296 : 800];   // per group (N+4)^2 padded |level| map (rate contexts)
  int16_t qc[NMAX == 32 ? 1024 : 512];                // per group n*n quantised levels in scan order
  PIX rec[NMAX == 32 ? 1024 : 512];              // per group n*n reconstruction
  int16_t ac[NMAX * NMAX];          // CFL: (luma << 3) - average of the block
  uint32_t taskD[6], taskR[6];
  uint16_t taskE[6];
};

struct TileCtx {
  const B200Frame *f;
  const B200RateTables *rt;
  const void *src[3];     // planes of PIX samples (uint8_t: 8-bit video, uint16_t: 10-bit), `stride` samples per row
  void *rec[3];
  B200Mi *mi;
  uint16_t *eob[3];
  int16_t *coef[3];
  int stride;
  int tr0, tr1, tc0, tc1;
  int tile_col;           // tile column index (reconstruction stage: stream bookkeeping)
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int txs_of(int n) { return n == 4 ? 0 : n == 8 ? 1 : n == 16 ? 2 : 3; }
__device__ __forceinline__ int scan_base(int n) { return n == 4 ? 0 : n == 8 ? 16 : 80; }
// position -> scan index (tables in global memory, L1 resident: the index diverges per lane)
template <int N> __device__ __forceinline__ int iscan_at(int pos) {
  if (N == 32) return __ldg(&dv_iscan32_g[pos]);
  return __ldg(&dv_iscan_g[scan_base(N) + pos]);
}
// scan indices of ROW l of an N x N block, packed: N <= 16: bytes of words w[0..N/4-1]; N == 32: halfwords of w[0..15].
// One vector load per row; SCAN_AT(c) extracts entry c with a compile-time shift when c is a constant.
template <int N> struct ScanRow { uint32_t w[N == 32 ? 16 : (N >= 4 ? N / 4 : 1)]; };
template <int N> __device__ __forceinline__ void scan_row_load(int l, ScanRow<N> &r) {
  if (N == 4) r.w[0] = __ldg(reinterpret_cast<const uint32_t *>(dv_iscan_g + l * 4));
  else if (N == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2 *>(dv_iscan_g + 16 + l * 8)); r.w[0] = v.x; r.w[1] = v.y; }
  else if (N == 16) { const uint4 v = __ldg(reinterpret_cast<const uint4 *>(dv_iscan_g + 80 + l * 16)); r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w; }
  else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(dv_iscan32_g + l * 32) + k);
      r.w[4 * k] = v.x; r.w[4 * k + 1] = v.y; r.w[4 * k + 2] = v.z; r.w[4 * k + 3] = v.w;
    }
  }
}
template <int N> __device__ __forceinline__ int scan_row_at(const ScanRow<N> &r, int c) {
  return N == 32 ? (int)((r.w[c >> 1] >> (16 * (c & 1))) & 0xffffu) : (int)((r.w[c >> 2] >> (8 * (c & 3))) & 0xffu);
}
// scan index of coefficient (row l, column c) [pos = l * N + c] for a transform class: default scan for the 2-D class, row by
// row for TX_CLASS_VERT (mrow), column by column for TX_CLASS_HORIZ (mcol): AV1 spec get_scan(), oracle scan_pos()
template <int N> __device__ __forceinline__ int scan_index(int txc, int pos, int l, int c) {
  return txc == 0 ? iscan_at<N>(pos) : (txc == 2 ? pos : c * N + l);
}
// scan indices of the FOUR coefficients (row l, columns c .. c + 3; c a multiple of 4) in one load: bytes (N <= 16) or
// halfwords (N == 32) of the returned pair; scan_quad_at() extracts entry c & 3
template <int N> __device__ __forceinline__ uint2 scan_quad_load(int l, int c) {
  if (N == 32) return __ldg(reinterpret_cast<const uint2 *>(dv_iscan32_g + l * 32 + c));
  return make_uint2(__ldg(reinterpret_cast<const uint32_t *>(dv_iscan_g + scan_base(N) + l * N + c)), 0u);
}
template <int N> __device__ __forceinline__ int scan_quad_at(const uint2 &q, int c) {
  if (N == 32) return (int)((((c & 2) ? q.y : q.x) >> (16 * (c & 1))) & 0xffffu);
  return (int)((q.x >> (8 * (c & 3))) & 0xffu);
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
template <class SH, class E>
__device__ __noinline__ void build_edges(const TileCtx &t, SH &sh, const E *r, int slot, int x, int y, int n, int avail) {
  typedef typename SH::Pix PIX;
  const B200Frame *f = t.f;
  const int st = t.stride, lane = lane_id();
  const int ha = avail & 1, hl = (avail >> 1) & 1, have_ar = (avail >> 2) & 1, have_bl = (avail >> 3) & 1;
  const int max_x = f->mi_cols * 4 - 1, max_y = f->mi_rows * 4 - 1;
  const int base = 1 << (f->bit_depth - 1);
  PIX *A = sh.above[slot] + EOFF, *L = sh.left[slot] + EOFF;
  for (int i = lane; i < 2 * n; i += 32) {
    PIX a, l;
    if (!ha && hl) a = r[y * st + x - 1];
    else if (!ha && !hl) a = (PIX)(base - 1);
    else {
      int xx = (i < n || have_ar) ? x + i : x + n - 1;
      if (xx > max_x) xx = max_x;
      a = r[(y - 1) * st + xx];
    }
    if (!hl && ha) l = r[(y - 1) * st + x];
    else if (!hl && !ha) l = (PIX)(base + 1);
    else {
      int yy = (i < n || have_bl) ? y + i : y + n - 1;
      if (yy > max_y) yy = max_y;
      l = r[yy * st + x - 1];
    }
    A[i] = a; L[i] = l;
  }
  if (lane == 0) {
    PIX tl;
    if (ha && hl) tl = r[(y - 1) * st + x - 1];
    else if (ha) tl = r[(y - 1) * st + x];
    else if (hl) tl = r[y * st + x - 1];
    else tl = (PIX)base;
    A[-1] = tl; L[-1] = tl;
  }
}

// ---- intra predictors (AV1 spec 7.11.2.x): value of pixel (i, j) of an n x n block ------------------
template <class E>
__device__ __noinline__ int dc_value(int n, const E *A, const E *L, int ha, int hl, int bd) {
  if (!ha && !hl) return 1 << (bd - 1);
  const int lg = 31 - __clz(n);          // n is a power of two: the spec's divisions are shifts
  int s = 0;
#pragma unroll 1
  for (int i = 0; i < n; i++) s += (ha ? A[i] : 0) + (hl ? L[i] : 0);
  return (ha && hl) ? (s + n) >> (lg + 1) : (s + (n >> 1)) >> lg;
}
// the same value computed by the whole warp (edges in shared memory): lane i adds sample i (n <= 32), one REDUX
template <class E>
__device__ __noinline__ int dc_value_warp(int n, const E *A, const E *L, int ha, int hl, int bd) {
  if (!ha && !hl) return 1 << (bd - 1);
  const int lg = 31 - __clz(n), i = lane_id();
  const unsigned s = __reduce_add_sync(FULL, i < n ? (unsigned)((ha ? A[i] : 0) + (hl ? L[i] : 0)) : 0u);
  return (ha && hl) ? (int)(s + n) >> (lg + 1) : (int)(s + (n >> 1)) >> lg;
}
// A, L point at index 0 of the edge arrays (index -1 valid).  dc: precomputed DC value (used for DC_PRED only).
template <class E>
__device__ __forceinline__ int pred_px(int mode, int delta, int n, int i, int j, const E *A, const E *L, int dc) {
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
    const E *e = A;
    if (base < -1) { idx = (i << 6) - (j + 1) * dy; base = idx >> 6; e = L; }
    const int shift = (idx >> 1) & 31;
    return (e[base] * (32 - shift) + e[base + 1] * shift + 16) >> 5;
  }
  // zones 1 (< 90, above row) and 3 (> 180, left column) are the same walk with the roles of i / j swapped
  const bool z1 = p_angle < 90;
  const E *e = z1 ? A : L;
  const int d = dv_dr_intra_derivative[z1 ? p_angle : 270 - p_angle], max_base = 2 * n - 1;
  const int idx = ((z1 ? i : j) + 1) * d, base = (idx >> 6) + (z1 ? j : i), shift = (idx >> 1) & 31;
  return base < max_base ? (e[base] * (32 - shift) + e[base + 1] * shift + 16) >> 5 : e[max_base];
}

// ---- 1-D transforms on register arrays -----------------------------------------------------
enum { K_DCT = 0, K_ADST = 1, K_IDT = 2 };
__device__ __forceinline__ void tx_kernels(int tx_type, int &vert, int &horz) {
  // V_DCT: DCT down the columns, identity along the rows; H_DCT: the other way round
  vert = (tx_type == TX_ADST_DCT || tx_type == TX_ADST_ADST) ? K_ADST : ((tx_type == TX_IDTX || tx_type == TX_H_DCT) ? K_IDT : K_DCT);
  horz = (tx_type == TX_DCT_ADST || tx_type == TX_ADST_ADST) ? K_ADST : ((tx_type == TX_IDTX || tx_type == TX_V_DCT) ? K_IDT : K_DCT);
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
// 8-bit video: every value this is applied to is below 2^30 in magnitude (DESIGN.md "int32 butterflies"), so the rounding
// add cannot overflow 32 bits; 10-bit keeps the 64-bit form
template <bool HBD> __device__ __forceinline__ int32_t rshift_round_t(int32_t v, int bit) {
  if (HBD) return rshift_round(v, bit);
  return bit <= 0 ? (int32_t)((uint32_t)v << (-bit)) : (v + (1 << (bit - 1))) >> bit;
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
// unroll factors of the per-coefficient quantisation / rate loops (A/B knobs: more ILP against a larger footprint)
#ifndef B200_UQ
#define B200_UQ 1
#endif
#ifndef B200_UR
#define B200_UR 1
#endif
#define B200_STR2(x) #x
#define B200_STR(x) B200_STR2(x)
#define B200_UNROLL(n) _Pragma(B200_STR(unroll n))

// RATE = false (reconstruction stage): the coefficient-rate model is skipped, TrialOut.rate is 0.
template <int N, bool HBD, class SH, bool RATE = true>
__device__ __noinline__ TrialOut trial_batch(const TileCtx &t, SH &sh, int plane, const typename SH::Pix *src, const typename SH::Pix *pred,
                                             int tx_type, int y_mode) {
  typedef typename SH::Pix PIX;
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
  PIX *rec = sh.rec + g * N2;
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
      for (int r = 0; r < N; r++) buf[r * S + l] = rshift_round_t<HBD>(o[r], shift1);
      __syncwarp();
    } else {
#pragma unroll
      for (int c = 0; c < N; c++) buf[l * S + c] = o[c];
    }
  }
  // ---- quantise / dequantise row l in place (oracle quantize()); exact division by reciprocal multiplication
  int eob = 0;
  const int txc = plane == 0 ? b200_tx_class(tx_type) : 0;      // transform class: scan order and coefficient contexts
  {
    // AC coefficients in the loop (one scan-table load per four of them); the DC coefficient (lane 0 of the group) after it
    const uint32_t dc_q = (uint32_t)f->dc_q, ac_q = (uint32_t)f->ac_q;
    const uint32_t ac_rcp = 0xffffffffu / ac_q, ac_off = (ac_q * 98) >> 8;
    uint2 sq = make_uint2(0u, 0u);
    B200_UNROLL(B200_UQ)
    for (int c = 0; c < N; c++) {
      const int pos = l * N + c;
      if ((c & 3) == 0 && txc == 0) sq = scan_quad_load<N>(l, c);
      if (pos == 0) continue;
      const int32_t cf = buf[l * S + c];
      const uint32_t a = ((uint32_t)(cf < 0 ? -cf : cf) << qshift) + ac_off;
      uint32_t lv = __umulhi(a, ac_rcp);        // floor(a / q) or one/two less
      uint32_t rem = a - lv * ac_q;
      if (rem >= ac_q) { lv++; rem -= ac_q; }
      if (rem >= ac_q) lv++;
      if (lv > 32767) lv = 32767;
      const int32_t dqa = (int32_t)((lv * ac_q) & 0xFFFFFF) >> qshift;
      const int si = txc == 0 ? scan_quad_at<N>(sq, c) : (txc == 2 ? pos : c * N + l);
      qc[si] = (int16_t)(cf < 0 ? -(int)lv : (int)lv);
      buf[l * S + c] = cf < 0 ? -dqa : dqa;
      lev[l * LS + c] = (uint8_t)(lv > 127 ? 127 : lv);
      if (lv) eob = max(eob, si + 1);
    }
    if (l == 0) {       // DC: scan index 0 in every class
      const uint32_t dc_rcp = 0xffffffffu / dc_q, dc_off = (dc_q * 109) >> 8;
      const int32_t cf = buf[0];
      const uint32_t a = ((uint32_t)(cf < 0 ? -cf : cf) << qshift) + dc_off;
      uint32_t lv = __umulhi(a, dc_rcp);
      uint32_t rem = a - lv * dc_q;
      if (rem >= dc_q) { lv++; rem -= dc_q; }
      if (rem >= dc_q) lv++;
      if (lv > 32767) lv = 32767;
      const int32_t dqa = (int32_t)((lv * dc_q) & 0xFFFFFF) >> qshift;
      qc[0] = (int16_t)(cf < 0 ? -(int)lv : (int)lv);
      buf[0] = cf < 0 ? -dqa : dqa;
      lev[0] = (uint8_t)(lv > 127 ? 127 : lv);
      if (lv) eob = max(eob, 1);
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
      rate += rt->eob_pt[txs][ptype][txc != 0][tg - 1];
      const int nb = tg < 3 ? 0 : tg - 2;
      if (nb > 0) {
        const int extra = eob - gstart;
        rate += rt->eob_extra[txs][ptype][tg - 3][(extra >> (nb - 1)) & 1];
        rate += 256u * (uint32_t)(nb - 1);
      }
    }
    uint2 sq = make_uint2(0u, 0u);
    B200_UNROLL(B200_UR)
    for (int c = 0; c < N; c++) {
      const int pos = l * N + c;
      if ((c & 3) == 0 && txc == 0) sq = scan_quad_load<N>(l, c);
      const int si = txc == 0 ? scan_quad_at<N>(sq, c) : (txc == 2 ? pos : c * N + l);
      if (si >= eob) continue;
      const int level = abs((int)qc[si]);
      const uint8_t *L = lev + l * LS + c;
      if (si == eob - 1) {
        const int c4 = si == 0 ? 0 : (si <= N2 / 8 ? 1 : (si <= N2 / 4 ? 2 : 3));
        rate += rt->coeff_base_eob[txs][ptype][c4][min(level, 3) - 1];
      } else {
        // neighbourhood and offset by transform class (AV1 spec Sig_Ref_Diff_Offset / Coeff_Base_Pos_Ctx_Offset; oracle coef_rate)
        int mag = min((int)L[1], 3) + min((int)L[LS], 3);
        if (txc == 0) mag += min((int)L[LS + 1], 3) + min((int)L[2], 3) + min((int)L[2 * LS], 3);
        else if (txc == 1) mag += min((int)L[2], 3) + min((int)L[3], 3) + min((int)L[4], 3);
        else mag += min((int)L[2 * LS], 3) + min((int)L[3 * LS], 3) + min((int)L[4 * LS], 3);
        int ctx = min((mag + 1) >> 1, 4);
        if (txc == 0) {
          if (l == 0 && c == 0) ctx = 0;
          else ctx += (l + c < 2) ? 1 : ((l + c < 4) ? 6 : 21);
        } else ctx += 26 + 5 * min(txc == 2 ? l : c, 2);
        rate += rt->coeff_base[txs][ptype][ctx][min(level, 3)];
      }
      if (level > 2) {
        int mag = min((int)L[1], 15) + min((int)L[LS], 15) + min((int)(txc == 0 ? L[LS + 1] : (txc == 1 ? L[2] : L[2 * LS])), 15);
        mag = min((mag + 1) >> 1, 6);
        const bool near = txc == 0 ? (l < 2 && c < 2) : (txc == 1 ? c == 0 : l == 0);
        const int bctx = (l == 0 && c == 0) ? mag : (near ? mag + 7 : mag + 14);
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
        for (int c = 0; c < N; c++) buf[l * S + c] = rshift_round_t<HBD>(o[c], row_shift);
      } else {
#pragma unroll
        for (int r = 0; r < N; r++) {
          int v = (int)pred[r * N + l] + rshift_round_t<HBD>(o[r], 4);
          v = v < 0 ? 0 : (v > mx ? mx : v);
          rec[r * N + l] = (PIX)v;
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
      rec[r * N + l] = (PIX)v;
      const int d = (int)src[r * N + l] - v;
      sse += (uint32_t)(d * d);
    }
    __syncwarp();
    __syncwarp();
  }
  sse = group_sum<N>(sse);
  if (N >= 8 && RATE && f->psy) {
    // Tune::Psychovisual (oracle trial_plane): luma is scored per 8x8 with the variance-aware error of records.h.  Lane l owns
    // column l of the reconstruction it just wrote; the five sums of an 8x8 are reduced over its 8 lanes.  Every lane takes the
    // same path (the shuffles are warp wide); groups that code a chroma plane keep their squared error.
    // Row group rg's five sums end up on every lane of an 8-lane segment; the lane with (l & 7) == rg keeps them, so that
    // after the loop the N/8 lanes (l & 7) < N/8 of a segment score the N/8 blocks of their column of 8x8 blocks at once.
    unsigned long long dist = 0;
    uint32_t kss = 0, ksd = 0, ks2 = 0, kd2 = 0, ksp = 0;
#pragma unroll 1
    for (int rg = 0; rg < N / 8; rg++) {
      uint32_t ss = 0, sd = 0, s2 = 0, d2 = 0, sp = 0;
#pragma unroll 1
      for (int r = rg * 8; r < rg * 8 + 8; r++) {
        const uint32_t s = src[r * N + l], d = rec[r * N + l];
        ss += s; sd += d; s2 += s * s; d2 += d * d; sp += s * d;
      }
#pragma unroll
      for (int m = 4; m > 0; m >>= 1) {
        ss += __shfl_xor_sync(FULL, ss, m); sd += __shfl_xor_sync(FULL, sd, m); s2 += __shfl_xor_sync(FULL, s2, m);
        d2 += __shfl_xor_sync(FULL, d2, m); sp += __shfl_xor_sync(FULL, sp, m);
      }
      if ((l & 7) == rg) { kss = ss; ksd = sd; ks2 = s2; kd2 = d2; ksp = sp; }
    }
    if ((l & 7) < N / 8) dist = b200_psy_dist8x8(kss, ksd, ks2, kd2, ksp, bd);
#pragma unroll
    for (int m = N / 2; m > 0; m >>= 1) dist += __shfl_xor_sync(FULL, dist, m);
    if (plane == 0) sse = dist > 0xffffffffull ? 0xffffffffu : (uint32_t)dist;
  }
  __syncwarp();
  TrialOut o;
  o.eob = eob; o.rate = rate; o.sse = sse;
  return o;
}

template <bool HBD, class SH>
__device__ __forceinline__ TrialOut trial_dispatch(int n, const TileCtx &t, SH &sh, int plane, const typename SH::Pix *src,
                                                   const typename SH::Pix *pred, int tx_type, int y_mode) {
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
template <class E>
__device__ __forceinline__ void ief_filter_edge(E *buf, int sz, int strength) {
  if (strength) {
    const int k0 = strength == 3 ? 2 : 0, k1 = strength == 1 ? 4 : (strength == 2 ? 5 : 4), k2 = strength == 1 ? 8 : (strength == 2 ? 6 : 4);
    E o[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const int i = 1 + lane_id() + 32 * r;
      int sum = 0;
      if (i < sz) {
        const int e0 = buf[max(i - 2, 0) - 1], e1 = buf[i - 2], e2 = buf[i - 1], e3 = buf[min(i + 1, sz - 1) - 1], e4 = buf[min(i + 2, sz - 1) - 1];
        sum = k0 * (e0 + e4) + k1 * (e1 + e3) + k2 * e2;
      }
      o[r] = (E)((sum + 8) >> 4);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; r++) { const int i = 1 + lane_id() + 32 * r; if (i < sz) buf[i - 1] = o[r]; }
  }
  __syncwarp();
}
// buf: index -2 .. 2 * num_px - 2 after the call, num_px <= 16
template <class E>
__device__ __forceinline__ void ief_upsample_edge(E *buf, int num_px, int bd) {
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
  if (i < num_px) { buf[2 * i - 1] = (E)v; buf[2 * i] = (E)m; }
  if (i == 0) buf[-2] = (E)corner;
  __syncwarp();
}
template <class E>
__device__ __forceinline__ int pred_dir_px(int p_angle, int n, int i, int j, const E *A, const E *L, int up_a, int up_l) {
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
template <class E>
__device__ __noinline__ void predict_block_ief(int mode, int delta, int lg, const E *above, const E *left, int ha, int hl,
                                               int bd, int ftype, E *ef_a, E *ef_l, E *dst) {
  const int n = 1 << lg, p_angle = dv_mode_to_angle[mode] + 3 * delta;
  E *A = ef_a + EOFF, *L = ef_l + EOFF;
#pragma unroll 1
  for (int i = lane_id() - 1; i < 2 * n; i += 32) { A[i] = above[EOFF + i]; L[i] = left[EOFF + i]; }
  __syncwarp();
  if (p_angle != 90 && p_angle != 180) {
    if (p_angle > 90 && p_angle < 180 && n >= 16) {
      const int c = (L[0] * 5 + A[-1] * 6 + A[0] * 5 + 8) >> 4;
      __syncwarp();
      if (lane_id() == 0) { A[-1] = (E)c; L[-1] = (E)c; }
      __syncwarp();
    }
    if (ha) ief_filter_edge(A, n + (p_angle < 90 ? n : 0) + 1, ief_strength(2 * n, ftype, p_angle - 90));
    if (hl) ief_filter_edge(L, n + (p_angle > 180 ? n : 0) + 1, ief_strength(2 * n, ftype, p_angle - 180));
  }
  const int up_a = ief_use_upsample(2 * n, ftype, p_angle - 90), up_l = ief_use_upsample(2 * n, ftype, p_angle - 180);
  if (up_a) ief_upsample_edge(A, n + (p_angle < 90 ? n : 0), bd);
  if (up_l) ief_upsample_edge(L, n + (p_angle > 180 ? n : 0), bd);
#pragma unroll 1
  for (int p = lane_id(); p < n * n; p += 32) dst[p] = (E)pred_dir_px(p_angle, n, p >> lg, p & (n - 1), A, L, up_a, up_l);
  __syncwarp();
}

// whole-block prediction into shared memory, pixels spread over the 32 lanes (the only place pred_px is expanded)
template <class E>
__device__ __noinline__ void predict_block(int mode, int delta, int lg, const E *above, const E *left, int ha, int hl, int bd,
                                           E *dst) {
  const E *A = above + EOFF, *L = left + EOFF;
  const int n = 1 << lg;
  const int dc = mode == DC_PRED ? dc_value_warp(n, A, L, ha, hl, bd) : 0;
#pragma unroll 1
  for (int p = lane_id(); p < n * n; p += 32) dst[p] = (E)pred_px(mode, delta, n, p >> lg, p & (n - 1), A, L, dc);
}

// ftype: -1 = intra edge filter off, 0 / 1 = on with that filter type
template <class SH>
__device__ __forceinline__ void predict_any(SH &sh, int ftype, int mode, int delta, int lg, const typename SH::Pix *above, const typename SH::Pix *left, int ha,
                                            int hl, int bd, typename SH::Pix *dst) {
  if (ftype >= 0 && mode >= V_PRED && mode <= D67_PRED) predict_block_ief(mode, delta, lg, above, left, ha, hl, bd, ftype, sh.ef[0], sh.ef[1], dst);
  else predict_block(mode, delta, lg, above, left, ha, hl, bd, dst);
}

// ---- chroma from luma (oracle cfl_ac / cfl_pick_alpha / cfl_predict), pixels spread over the 32 lanes
template <class E>
__device__ __noinline__ void cfl_build_ac(const E *luma, int lg, int16_t *ac) {
  const int N2 = 1 << (2 * lg);
  long long sum = 0;
#pragma unroll 1
  for (int p = lane_id(); p < N2; p += 32) sum += luma[p] << 3;
  const int avg = (int)((warp_sum_ll(sum) + (N2 >> 1)) >> (2 * lg));
#pragma unroll 1
  for (int p = lane_id(); p < N2; p += 32) ac[p] = (int16_t)((luma[p] << 3) - avg);
  __syncwarp();
}
template <class E>
__device__ __noinline__ int cfl_pick_alpha(const int16_t *ac, const E *src, int N2, int dc) {
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
template <class E>
__device__ __noinline__ void cfl_predict_block(const int16_t *ac, int N2, int dc, int alpha, int bd, E *dst) {
  const int mx = (1 << bd) - 1;
#pragma unroll 1
  for (int p = lane_id(); p < N2; p += 32) {
    const int sc = alpha * ac[p];
    const int v = dc + (sc < 0 ? -((-sc + 32) >> 6) : ((sc + 32) >> 6));
    dst[p] = (E)(v < 0 ? 0 : (v > mx ? mx : v));
  }
}
__device__ __noinline__ uint32_t cfl_rate(const B200RateTables *rt, int au, int av) {
  const int su = au == 0 ? 0 : (au < 0 ? 1 : 2), sv = av == 0 ? 0 : (av < 0 ? 1 : 2);
  uint32_t r = rt->cfl_sign[su * 3 + sv - 1];
  if (su) r += rt->cfl_alpha[(su - 1) * 3 + sv][abs(au) - 1];
  if (sv) r += rt->cfl_alpha[(sv - 1) * 3 + su][abs(av) - 1];
  return r;
}

// top-4 of 13 mode costs held by lanes 0..12 (kInf when not allowed): lowest cost first, ties -> lowest index; -1 = none
__device__ __noinline__ void rank_top4(long long c, int &m0, int &m1, int &m2, int &m3) {
  const int lane = lane_id();
  unsigned taken = 0;
  int m[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    long long j = (lane < 13 && !((taken >> lane) & 1)) ? c : kInf;
    int i = lane & 15;
    warp_argmin(j, i);
    m[r] = j == kInf ? -1 : i;
    if (m[r] >= 0) taken |= 1u << m[r];
  }
  m0 = m[0]; m1 = m[1]; m2 = m[2]; m3 = m[3];
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
  t.mi = job.mi; t.stride = job.stride;
}

// ---- vectorised staging of a source block: global -> registers -> shared memory --------------------------------------
// source block (n x n at frame position x, y) -> registers: one row segment of up to 16 bytes per lane
template <class PIX>
struct SrcRegs { uint4 v; };
template <class PIX>
__device__ __forceinline__ void src_fetch(const PIX *plane, int stride, int x, int y, int n, SrcRegs<PIX> &o, int half /*u16 n=32 second half etc.*/) {
  // bytes per row = n * sizeof(PIX); segments of 16 bytes (or the whole row when shorter)
  const int row_bytes = n * (int)sizeof(PIX), seg = row_bytes < 16 ? row_bytes : 16, segs_per_row = row_bytes / seg;
  const int idx = lane_id() + 32 * half, r = idx / segs_per_row, s = idx % segs_per_row;
  o.v = make_uint4(0, 0, 0, 0);
  if (r >= n) return;
  const char *p = reinterpret_cast<const char *>(plane + (size_t)(y + r) * stride + x) + s * seg;
  if (seg == 16) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) o.v = *reinterpret_cast<const uint4 *>(p);
    else { const uint2 a = *reinterpret_cast<const uint2 *>(p), b = *reinterpret_cast<const uint2 *>(p + 8); o.v = make_uint4(a.x, a.y, b.x, b.y); }
  } else if (seg == 8) { const uint2 a = *reinterpret_cast<const uint2 *>(p); o.v.x = a.x; o.v.y = a.y; }
  else o.v.x = *reinterpret_cast<const uint32_t *>(p);
}
template <class PIX>
__device__ __forceinline__ void src_store(PIX *dst /* n x n row-major, 16-byte aligned */, int n, const SrcRegs<PIX> &o, int half) {
  const int row_bytes = n * (int)sizeof(PIX), seg = row_bytes < 16 ? row_bytes : 16, segs_per_row = row_bytes / seg;
  const int idx = lane_id() + 32 * half, r = idx / segs_per_row, s = idx % segs_per_row;
  if (r >= n) return;
  char *p = reinterpret_cast<char *>(dst) + r * row_bytes + s * seg;
  if (seg == 16) *reinterpret_cast<uint4 *>(p) = o.v;
  else if (seg == 8) *reinterpret_cast<uint2 *>(p) = make_uint2(o.v.x, o.v.y);
  else *reinterpret_cast<uint32_t *>(p) = o.v.x;
}
// number of 32-lane fetch rounds a block of n x n samples needs
template <class PIX> __device__ __forceinline__ int src_rounds(int n) {
  const int row_bytes = n * (int)sizeof(PIX), seg = row_bytes < 16 ? row_bytes : 16;
  return (n * (row_bytes / seg) + 31) / 32;
}


}  // namespace
#endif  // B200_AV1_DEV_CUH_
