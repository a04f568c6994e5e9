// AV1 in-loop filter stage on the device: deblocking (spec 7.14) and CDEF (spec 7.15) with the
// per-superblock CDEF strength search (squared error against the source).  Streaming kernels over
// the reconstruction written by the RDO kernel; every job (frame) of the batch in one launch.
//
// Reference: rav1e deblock.rs / cdef.rs / rdo_loop_decision [EXT: crates.io rav1e 0.8.1], switched on
// by ravif's SpeedTweaks (cdef / lrf flags, /root/reference/ravif/src/av1encoder.rs:589-590).
// Bit-exact CPU restatement: oracle/av1_filters.c; the normative arithmetic is pinned by dav1d.
#ifdef B200_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "records.h"
#include "rdo.h"
#ifndef B200_EMU
#include <mutex>
#endif
#ifdef B200_EMU
#define B200_MARK_F(name, st) ((void)0)
#else
#define B200_MARK_F(name, st) b200_mark(name, (void *)(st))
#endif

#define AV1_TABLE(type, name, dims) __constant__ const type name dims
#include "sgr_tables_dev.h"
#undef AV1_TABLE

namespace {

// slot of frame row y in job.dbrows (the deblocked rows 54..57 of every superblock row, kept for the loop restoration)
__device__ __forceinline__ int db_row_slot(int y) { return (y >> 6) * 4 + ((y & 63) - 54); }
__device__ __forceinline__ int iabs_d(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int clip3_d(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// one edge sample; px -> q0, step = tap distance
template <class PIX>
__device__ void deblock_sample(PIX *px, int step, int filter_len, int lvl, int sharp, int bd) {
  const int shift = sharp > 4 ? 2 : (sharp > 0 ? 1 : 0);
  int limit = sharp > 0 ? clip3_d(1, 9 - sharp, lvl >> shift) : max(1, lvl >> shift);
  const int blimit = (2 * (lvl + 2) + limit) << (bd - 8);
  const int thresh = (lvl >> 4) << (bd - 8);
  limit <<= (bd - 8);
  const int one = 1 << (bd - 8);
  int p[7], q[7];
  const int taps = filter_len == 4 ? 2 : (filter_len == 6 ? 3 : (filter_len == 8 ? 4 : 7));
#pragma unroll
  for (int i = 0; i < 7; i++)
    if (i < taps) { q[i] = px[i * step]; p[i] = px[-(i + 1) * step]; } else { q[i] = 0; p[i] = 0; }
  const int hev = iabs_d(p[1] - p[0]) > thresh || iabs_d(q[1] - q[0]) > thresh;
  int mask = iabs_d(p[1] - p[0]) <= limit && iabs_d(q[1] - q[0]) <= limit && (iabs_d(p[0] - q[0]) * 2 + iabs_d(p[1] - q[1]) / 2) <= blimit;
  if (filter_len >= 6) mask = mask && iabs_d(p[2] - p[1]) <= limit && iabs_d(q[2] - q[1]) <= limit;
  if (filter_len >= 8) mask = mask && iabs_d(p[3] - p[2]) <= limit && iabs_d(q[3] - q[2]) <= limit;
  if (!mask) return;
  int flat = 0, flat2 = 0;
  if (filter_len >= 6) {
    flat = iabs_d(p[1] - p[0]) <= one && iabs_d(q[1] - q[0]) <= one && iabs_d(p[2] - p[0]) <= one && iabs_d(q[2] - q[0]) <= one;
    if (filter_len >= 8) flat = flat && iabs_d(p[3] - p[0]) <= one && iabs_d(q[3] - q[0]) <= one;
  }
  if (filter_len == 16)
    flat2 = iabs_d(p[4] - p[0]) <= one && iabs_d(q[4] - q[0]) <= one && iabs_d(p[5] - p[0]) <= one && iabs_d(q[5] - q[0]) <= one &&
            iabs_d(p[6] - p[0]) <= one && iabs_d(q[6] - q[0]) <= one;
  if (filter_len == 4 || !flat) {
    const int lo = -(1 << (bd - 1)), hi = (1 << (bd - 1)) - 1, off = 0x80 << (bd - 8);
    const int ps1 = p[1] - off, ps0 = p[0] - off, qs0 = q[0] - off, qs1 = q[1] - off;
    int f = hev ? clip3_d(lo, hi, ps1 - qs1) : 0;
    f = clip3_d(lo, hi, f + 3 * (qs0 - ps0));
    const int f1 = clip3_d(lo, hi, f + 4) >> 3, f2 = clip3_d(lo, hi, f + 3) >> 3;
    px[0] = (PIX)(clip3_d(lo, hi, qs0 - f1) + off);
    px[-step] = (PIX)(clip3_d(lo, hi, ps0 + f2) + off);
    if (!hev) {
      const int f3 = (f1 + 1) >> 1;
      px[step] = (PIX)(clip3_d(lo, hi, qs1 - f3) + off);
      px[-2 * step] = (PIX)(clip3_d(lo, hi, ps1 + f3) + off);
    }
    return;
  }
  int n, lg;
  if (filter_len == 16 && flat2) { n = 6; lg = 4; }
  else if (filter_len == 6) { n = 2; lg = 3; }
  else { n = 3; lg = 3; }
  // Output i (-n <= i < n) = (sum_{j=-n..n} tap_j * s[clip(i + j)] + round) >> lg with taps 1 except 2 at j = 0 (n == 3) or at
  // |j| < 2 (n == 2, 6); s = the 2n + 2 samples around the edge, indices clipped to [-(n + 1), n] (AV1 spec 7.14.6.3/4).  The
  // all-ones part is a window sum that slides by one sample per output; the doubled taps never need clipping.
  int s[14];
  for (int i = 0; i <= n; i++) { s[n - i] = p[i]; s[n + 1 + i] = q[i]; }
  const int *sc = s + n + 1;          // sc[k], k = -(n + 1) .. n
  int win = 0;
  for (int j = -n; j <= n; j++) win += sc[clip3_d(-(n + 1), n, -n + j)];
  const int rnd = 1 << (lg - 1);
  for (int i = -n; i < n; i++) {
    const int t = win + (n == 3 ? sc[i] : sc[i - 1] + sc[i] + sc[i + 1]);
    px[i * step] = (PIX)((t + rnd) >> lg);
    win += sc[min(i + 1 + n, n)] - sc[max(i - n, -(n + 1))];
  }
}

}  // namespace

namespace {
__constant__ const int8_t dv_cdef_dir[8][2][2] = {{{-1, 1}, {-2, 2}}, {{0, 1}, {-1, 2}}, {{0, 1}, {0, 2}}, {{0, 1}, {1, 2}},
                                                  {{1, 1}, {2, 2}},   {{1, 0}, {2, 1}},  {{1, 0}, {2, 0}}, {{1, 0}, {2, -1}}};
__device__ __forceinline__ int floor_log2_d(unsigned v) { return 31 - __clz((int)v); }
__device__ __forceinline__ int constrain_d(int diff, int threshold, int damping) {
  if (!threshold) return 0;
  const int adj = max(0, damping - floor_log2_d((unsigned)threshold));
  const int mag = iabs_d(diff);
  const int v = clip3_d(0, mag, threshold - (mag >> adj));
  return diff < 0 ? -v : v;
}

__device__ void cdef_direction(const uint16_t *img, int stride, int bd, int *dir_out, int *var_out) {
  int cost[8], partial[8][15];
  for (int d = 0; d < 8; d++) { cost[d] = 0; for (int k = 0; k < 15; k++) partial[d][k] = 0; }
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) {
      const int x = (img[i * stride + j] >> (bd - 8)) - 128;
      partial[0][i + j] += x;
      partial[1][i + j / 2] += x;
      partial[2][i] += x;
      partial[3][3 + i - j / 2] += x;
      partial[4][7 + i - j] += x;
      partial[5][3 - i / 2 + j] += x;
      partial[6][j] += x;
      partial[7][i / 2 + j] += x;
    }
  const int div_table[9] = {0, 840, 420, 280, 210, 168, 140, 120, 105};
  for (int i = 0; i < 8; i++) { cost[2] += partial[2][i] * partial[2][i]; cost[6] += partial[6][i] * partial[6][i]; }
  cost[2] *= div_table[8]; cost[6] *= div_table[8];
  for (int i = 0; i < 7; i++) {
    cost[0] += (partial[0][i] * partial[0][i] + partial[0][14 - i] * partial[0][14 - i]) * div_table[i + 1];
    cost[4] += (partial[4][i] * partial[4][i] + partial[4][14 - i] * partial[4][14 - i]) * div_table[i + 1];
  }
  cost[0] += partial[0][7] * partial[0][7] * div_table[8];
  cost[4] += partial[4][7] * partial[4][7] * div_table[8];
  for (int i = 1; i < 8; i += 2) {
    for (int j = 0; j < 5; j++) cost[i] += partial[i][3 + j] * partial[i][3 + j];
    cost[i] *= div_table[8];
    for (int j = 0; j < 3; j++) cost[i] += (partial[i][j] * partial[i][j] + partial[i][10 - j] * partial[i][10 - j]) * div_table[2 * j + 2];
  }
  int best = 0, bc = 0;
  for (int d = 0; d < 8; d++) if (cost[d] > bc) { bc = cost[d]; best = d; }
  *dir_out = best;
  *var_out = (bc - cost[(best + 4) & 7]) >> 10;
}

// CDEF output of one pixel of a non-skip 8x8 block for strength pair (pri, sec) (AV1 spec 7.15.2), read from a
// staged (64 + 4)^2 tile of the superblock (kCdefTW samples per row, 2-sample border, kCdefNA marks
// samples outside the frame): lx, ly = position inside the superblock.
constexpr int kFH = 7;                   // halo of the staged region: the deblocking filters reach 7 samples past an edge
constexpr int kFW = 64 + 2 * kFH;        // staged rows / columns
constexpr int kCdefTW = 80;              // row stride of the staged region in shared memory
constexpr uint16_t kCdefNA = 0xFFFF;
#ifndef B200_EMU
// ---- TMA (cp.async.bulk.tensor) + mbarrier: the tensor-memory-access unit moves the (64 + 14)^2 region of a plane from
// HBM / L2 into shared memory on its own - no address arithmetic, no load instructions, no registers of the SM - and signals
// an mbarrier with the byte count when the tile has landed.  Out-of-frame parts of the box arrive as zeros.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void *bar, uint32_t phase) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return ok != 0;
}
// A tensor map that lives in GLOBAL memory is read through the tensor-map proxy, which caches descriptors by address: the
// arena re-uses addresses from call to call with new contents, so the issuing thread first acquires the 128 bytes.
__device__ __forceinline__ void tmap_acquire(const void *tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int x, int y, void *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
#endif
#ifndef B200_EMU
__device__ unsigned int b200_tma_timeouts;      // TMA tiles that did not land in time (diagnostic; read by b200_filters_tma_timeouts)
#endif
// Box of the tensor map.  The byte offset of the box inside a row must be a multiple of 16 (an 8-byte-aligned start is an
// illegal instruction at the UTMALDG, found with compute-sanitizer): 16-bit samples start at x0 - 8 (80 wide, 160 bytes),
// 8-bit samples at x0 - 16 (96 wide); 78 rows from y0 - 7 in both cases.
constexpr int kTmaBoxH = kFW;
constexpr int kDbPlaneBytes = ((kFW * kCdefTW * 2 + 127) / 128) * 128;      // one plane of the fused kernel's shared memory (working copy + landing zone)
template <class PIX> struct TmaBox { static constexpr int kOff = sizeof(PIX) == 1 ? 16 : 8, kW = sizeof(PIX) == 1 ? 96 : 80; };
__device__ __forceinline__ int cdef_pixel_tile(const uint16_t *T, int lx, int ly, int pri, int sec, int dir, int damping, int cshift) {
  const uint16_t *c = T + (ly + kFH) * kCdefTW + lx + kFH;
  const int x = *c;
  if (!pri && !sec) return x;
  const int pt0 = ((pri >> cshift) & 1) ? 3 : 4, pt1 = ((pri >> cshift) & 1) ? 3 : 2;
  int sum = 0, mx = x, mn = x;
#pragma unroll
  for (int k = 0; k < 2; k++)
#pragma unroll
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      {
        const int v = c[sgn * (dv_cdef_dir[dir][k][0] * kCdefTW + dv_cdef_dir[dir][k][1])];
        if (v != kCdefNA) {
          sum += (k == 0 ? pt0 : pt1) * constrain_d(v - x, pri, damping);
          mx = max(mx, v); mn = min(mn, v);
        }
      }
#pragma unroll
      for (int doff = -2; doff <= 2; doff += 4) {
        const int d2 = (dir + doff) & 7;
        const int v = c[sgn * (dv_cdef_dir[d2][k][0] * kCdefTW + dv_cdef_dir[d2][k][1])];
        if (v != kCdefNA) {
          sum += (k == 0 ? 2 : 1) * constrain_d(v - x, sec, damping);
          mx = max(mx, v); mn = min(mn, v);
        }
      }
    }
  return clip3_d(mn, mx, x + ((8 + sum - (sum < 0)) >> 4));
}
// The strength search evaluates several (primary, secondary) strength pairs on the SAME taps: the 4 primary and 8 secondary
// tap samples of a pixel depend only on the block's direction.  They are loaded once; a tap outside the frame is replaced by
// the centre sample (difference 0, no effect on the clipping range - the same as skipping it, spec 7.15.2).
struct CdefTaps { int x, p[4], s[8]; };
__device__ __forceinline__ void cdef_load_taps(const uint16_t *c, const int16_t *off /* [8][2]: dy * stride + dx */, int dir, CdefTaps &t) {
  t.x = *c;
  const int da = (dir - 2) & 7, db = (dir + 2) & 7;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int op = off[dir * 2 + k], oa = off[da * 2 + k], ob = off[db * 2 + k];
    int v;
    v = c[-op]; t.p[2 * k] = v == kCdefNA ? t.x : v;
    v = c[op]; t.p[2 * k + 1] = v == kCdefNA ? t.x : v;
    v = c[-oa]; t.s[4 * k] = v == kCdefNA ? t.x : v;
    v = c[oa]; t.s[4 * k + 1] = v == kCdefNA ? t.x : v;
    v = c[-ob]; t.s[4 * k + 2] = v == kCdefNA ? t.x : v;
    v = c[ob]; t.s[4 * k + 3] = v == kCdefNA ? t.x : v;
  }
}
__device__ __forceinline__ int constrain_adj(int diff, int threshold, int adj) {
  const int mag = iabs_d(diff);
  const int v = max(0, min(mag, threshold - (mag >> adj)));
  return diff < 0 ? -v : v;
}
// One strength pair on loaded taps = cdef_finish(cdef_pri_sum + cdef_sec_sum).  With only one of the two strengths the
// result provably stays inside the range of the taps it used (total tap weight 12/16), so the spec's clip to [min, max] is
// the identity and is skipped (`both` = false).
// weighted sum of the primary / secondary taps (pri, sec != 0); the secondary sum depends on the secondary strength only,
// so candidate pairs that share it (the chroma sets do) share the sum
__device__ __forceinline__ int cdef_pri_sum(const CdefTaps &t, int pri, int damping, int cshift) {
  const int x = t.x, adj = max(0, damping - floor_log2_d((unsigned)pri));
  const int w0 = ((pri >> cshift) & 1) ? 3 : 4, w1 = ((pri >> cshift) & 1) ? 3 : 2;
  return w0 * (constrain_adj(t.p[0] - x, pri, adj) + constrain_adj(t.p[1] - x, pri, adj)) +
         w1 * (constrain_adj(t.p[2] - x, pri, adj) + constrain_adj(t.p[3] - x, pri, adj));
}
__device__ __forceinline__ int cdef_sec_sum(const CdefTaps &t, int sec, int damping) {
  const int x = t.x, adj = max(0, damping - floor_log2_d((unsigned)sec));
  int s0 = 0, s1 = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) { s0 += constrain_adj(t.s[i] - x, sec, adj); s1 += constrain_adj(t.s[4 + i] - x, sec, adj); }
  return 2 * s0 + s1;
}
// sum: cdef_pri_sum (pri != 0) + cdef_sec_sum (sec != 0)
__device__ __forceinline__ int cdef_finish(const CdefTaps &t, int sum, bool both) {
  const int x = t.x;
  int y = x + ((8 + sum - (sum < 0)) >> 4);
  if (both) {
    int mn = x, mx = x;
#pragma unroll
    for (int i = 0; i < 4; i++) { mn = min(mn, t.p[i]); mx = max(mx, t.p[i]); }
#pragma unroll
    for (int i = 0; i < 8; i++) { mn = min(mn, t.s[i]); mx = max(mx, t.s[i]); }
    y = clip3_d(mn, mx, y);
  }
  return y;
}
// ---- loop restoration: self-guided filter parameter search (oracle/av1_restoration.c orc_lr_search_plane) ----------
// One CTA (256 threads) per (restoration unit, plane); grid.y = job, grid.z = plane.  The unit is walked stripe by
// stripe and in bands of 16 rows: a band's samples (CDEF output inside the stripe, deblocked rows outside it, AV1
// spec 7.17.6) are staged once in shared memory and filtered with every candidate parameter set; the per-set normal
// equations of the projection are accumulated in shared 64-bit counters, then one thread solves, quantises and scores
// each set (quadratic form, no second filter pass).  When job.lr_out is set (tests), the winner is applied as well.
constexpr int kLrBandH = 16, kLrMaxW = 96;
constexpr int kLrSW = kLrMaxW + 6, kLrSH = kLrBandH + 6;        // staged samples: 3-sample border
constexpr int kLrAW = kLrMaxW + 2, kLrAH = kLrBandH + 2;        // A / B: 1-entry border

struct LrShared {
  uint16_t S[kLrSH * kLrSW];
  uint16_t A[kLrAH * kLrAW];
  int32_t B[kLrAH * kLrAW];
  uint16_t F0[kLrBandH * kLrMaxW];
  uint16_t SRC[kLrBandH * kLrMaxW];     // source samples of the band (read once per band instead of once per candidate set)
  // raw box sums of the band, shared by every candidate set: 3x3 at all A/B rows, 5x5 at the odd rows only
  uint32_t R3a[kLrAH * kLrAW], R5a[((kLrAH + 1) / 2) * kLrAW];
  uint16_t R3b[kLrAH * kLrAW], R5b[((kLrAH + 1) / 2) * kLrAW];
  uint16_t xbx[256];      // x_by_xplus1 (spec 7.17.3): indexed by the per-lane variance, so a shared-memory copy (a __constant__ read serialises over distinct addresses)
  unsigned long long acc[8][B200_LR_CANDIDATES_FULL][6];  // per warp and candidate set: H00, H01, H11, C0, C1 (two's complement), S - one writer per slot, summed at the end
  B200LrUnit best;
};

// floor(i / d) by reciprocal multiplication, magic = ceil(2^20 / d): exact while i < 2^20 / d (here i < 22 * d, d <= 102)
struct LrDiv { uint32_t d, magic; };
__device__ __forceinline__ LrDiv lr_div_make(int d) { LrDiv r; r.d = (uint32_t)d; r.magic = ((1u << 20) + (uint32_t)d - 1) / (uint32_t)d; return r; }
__device__ __forceinline__ void lr_divmod(int i, const LrDiv &dv, int &q, int &r) {
  q = (int)(((uint32_t)i * dv.magic) >> 20);
  r = i - q * (int)dv.d;
}
__device__ __forceinline__ long long lr_round2(long long v, int n) { return n == 0 ? v : ((v + (1ll << (n - 1))) >> n); }

// box sums of radius r around staged sample (sy, sx)
__device__ __forceinline__ void lr_box(const uint16_t *S, int sy, int sx, int r, uint32_t *a_out, uint32_t *b_out) {
  uint32_t a = 0, b = 0;
  for (int dy = -r; dy <= r; dy++)
    for (int dx = -r; dx <= r; dx++) {
      const uint32_t v = S[(sy + dy) * kLrSW + sx + dx];
      a += v * v; b += v;
    }
  *a_out = a; *b_out = b;
}
// box sums -> a2 (A entry) and B entry for one parameter set (AV1 spec 7.17.4)
// B8 (8-bit samples): no pre-scaling, and a * n, d * d <= 25 * 25 * 255^2 fit 32 bits
template <bool B8>
__device__ __forceinline__ void lr_ab(const uint16_t *xbx, uint32_t a, uint32_t b, int bd, uint32_t s, uint32_t one_by_n, int n, uint16_t *a_out, int32_t *b_out) {
  // a <= 25 * 1023^2 < 2^25 and b <= 25 * 1023 at 10 bits: the pre-scaled values are in the 8-bit range, so 32 bits hold
  // a_s * n and d * d for either depth (bit depths above 10 are not produced by ravif)
  const int sh2 = B8 ? 0 : bd - 8;
  const uint32_t a_s = B8 ? a : (a + ((1u << (2 * sh2)) >> 1)) >> (2 * sh2), d = B8 ? b : (b + ((1u << sh2) >> 1)) >> sh2;
  const int32_t pp = (int32_t)(a_s * (uint32_t)n) - (int32_t)(d * d);
  const unsigned long long p = pp < 0 ? 0u : (uint32_t)pp;
  const unsigned long long z = (p * s + (1u << 19)) >> 20;
  const uint32_t a2 = xbx[z > 255 ? 255 : (int)z];
  *a_out = (uint16_t)a2;
  *b_out = (int32_t)(((256u - a2) * b * one_by_n + (1u << 11)) >> 12);      // (256 - a2) * b * one_by_n < 2^31 for r = 1 and r = 2
}

__device__ __forceinline__ long long lr_round_div(long long a, long long b) { return a >= 0 ? (a + b / 2) / b : -((-a + b / 2) / b); }
__device__ __forceinline__ int lr_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ int lr_subexp_bits(int lo, int hi, int ref, int v) {      // oracle subexp_bits(), k = 4
  const int mx = hi - lo, r = ref - lo, x = v - lo;
  int t;
  if ((r << 1) <= mx) t = x > 2 * r ? x : (x >= r ? (x - r) << 1 : ((r - x) << 1) - 1);
  else { const int r2 = mx - 1 - r, x2 = mx - 1 - x; t = x2 > 2 * r2 ? x2 : (x2 >= r2 ? (x2 - r2) << 1 : ((r2 - x2) << 1) - 1); }
  int i = 0, mk = 0, bits = 0;
  for (;;) {
    const int b2 = i ? 4 + i - 1 : 4, a = 1 << b2;
    if (mx <= mk + 3 * a) {
      const int nsym = mx - mk;
      int w = 0; while ((1 << w) <= nsym) w++;
      const int m = (1 << w) - nsym;
      return bits + ((t - mk) < m ? w - 1 : w);
    }
    bits++;
    if (t >= mk + a) { i++; mk += a; } else return bits + b2;
  }
}
}  // namespace

#ifdef B200_EMU
static int g_f_plane;
#endif
template <class PIX>
__global__ void __launch_bounds__(256) b200_lr_kernel(const B200RdoJob *jobs) {
  __shared__ LrShared sh;
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
#ifdef B200_EMU
  const int plane = g_f_plane;       // the emulator launches one plane at a time
#else
  const int plane = blockIdx.z;
#endif
  if (plane >= f->planes || !f->lr_type[plane]) return;
  const int nur = b200_lr_units(f->height), nuc = b200_lr_units(f->width);
  if ((int)blockIdx.x >= nur * nuc) return;
  const int ur = blockIdx.x / nuc, uc = blockIdx.x % nuc;
  const int stride = job.stride, W = f->width, H = f->height, bd = f->bit_depth;
  const int y0 = max(ur * 64 - 8, 0), y1 = ur == nur - 1 ? H : ur * 64 + 56;
  const int x0 = uc * 64, x1 = uc == nuc - 1 ? W : uc * 64 + 64, w = x1 - x0;
  // deb: the deblocked rows 54..57 of every superblock row kept by the fused kernel (the only ones read: spec 7.17.6)
  const PIX *deb = (const PIX *)job.dbrows[plane], *cdef = (const PIX *)job.fin[plane], *src = (const PIX *)job.src[plane];
  PIX *lr_out = (PIX *)job.lr_out[plane];
  const int tid = threadIdx.x;
  // the loops below walk (rows x columns) index spaces of at most 22 x 102 entries: row / column by reciprocal multiplication
  const LrDiv dv_w = lr_div_make(w), dv_w2 = lr_div_make(w + 2), dv_w6 = lr_div_make(w + 6);
  for (int i = tid; i < 8 * B200_LR_CANDIDATES_FULL * 6; i += 256) (&sh.acc[0][0][0])[i] = 0;
  sh.xbx[tid] = dv_x_by_xplus1[tid];
  __syncthreads();
  const int8_t cand[B200_LR_CANDIDATES] = {B200_LR_SETS};
  // sgr_complexity_full (ravif SpeedTweaks, av1encoder.rs:573): every one of the 16 parameter sets, else four
  const int n_sets = f->lr_full ? B200_LR_CANDIDATES_FULL : B200_LR_CANDIDATES;
#define LR_SET_OF(ci) (f->lr_full ? (ci) : (int)cand[(ci) & 3])
  // pass over the unit: phase 0 accumulates the normal equations of every candidate set, phase 1 (tests only) applies
  // the winner
  for (int phase = 0; phase < 2; phase++) {
    if (phase == 1 && !lr_out) break;
    for (int ys = y0; ys < y1;) {
      const int stripe = (ys + 8) >> 6, s_start = stripe * 64 - 8, s_end = stripe * 64 + 55;
      const int ye = min(min(s_end + 1, y1), ys + kLrBandH), bh = ye - ys;
      // stage the samples (AV1 spec 7.17.6 get_source_sample)
      for (int i = tid; i < (bh + 6) * (w + 6); i += 256) {
        int sy, sx;
        lr_divmod(i, dv_w6, sy, sx);
        int y = ys + sy - 3, x = x0 + sx - 3;
        x = min(max(x, 0), W - 1); y = min(max(y, 0), H - 1);
        uint16_t v;
        if (y < s_start) v = deb[(size_t)db_row_slot(max(y, s_start - 2)) * stride + x];
        else if (y > s_end) v = deb[(size_t)db_row_slot(min(y, s_end + 2)) * stride + x];
        else v = cdef[y * stride + x];
        sh.S[sy * kLrSW + sx] = v;
      }
      if (phase == 0)
        for (int i = tid; i < bh * w; i += 256) {
          int py, px;
          lr_divmod(i, dv_w, py, px);
          sh.SRC[py * kLrMaxW + px] = src[(ys + py) * stride + x0 + px];
        }
      __syncthreads();
      // raw box sums, once per band
      for (int i = tid; i < (bh + 2) * (w + 2); i += 256) {
        int ai, aj;
        lr_divmod(i, dv_w2, ai, aj);
        uint32_t a, b;
        lr_box(sh.S, ai + 2, aj + 2, 1, &a, &b);
        sh.R3a[ai * kLrAW + aj] = a; sh.R3b[ai * kLrAW + aj] = (uint16_t)b;
      }
      for (int i = tid; i < ((bh + 3) / 2) * (w + 2); i += 256) {
        int ri, aj;                                            // A row index 2 * ri <-> i = -1, 1, 3, ...
        lr_divmod(i, dv_w2, ri, aj);
        uint32_t a, b;
        lr_box(sh.S, 2 * ri + 2, aj + 2, 2, &a, &b);
        sh.R5a[ri * kLrAW + aj] = a; sh.R5b[ri * kLrAW + aj] = (uint16_t)b;
      }
      __syncthreads();
      const int ncand = phase == 0 ? n_sets : 1;
      for (int ci = 0; ci < ncand; ci++) {
        const int set = phase == 0 ? LR_SET_OF(ci) : sh.best.set;
        if (phase == 1 && !sh.best.use) {       // restoration off in this unit: the CDEF samples pass through
          for (int i = tid; i < bh * w; i += 256) {
            int py, px;
            lr_divmod(i, dv_w, py, px);
            lr_out[(ys + py) * stride + x0 + px] = (PIX)sh.S[(py + 3) * kLrSW + px + 3];
          }
          continue;
        }
        const int r0 = dv_sgr_params[set][0], r1 = dv_sgr_params[set][1];
        // per-thread partial sums of the band: at most ceil(16 * 95 / 256) = 6 pixels, every product below 2^28 in
        // magnitude (|f - u|, |src - u| < 2^14 at 10 bits) -> 32-bit accumulators, widened at the warp reduction
        int32_t h00 = 0, h01 = 0, h11 = 0, c0 = 0, c1 = 0, ss = 0;
        for (int pass = 0; pass < 2; pass++) {
          const int r = pass == 0 ? r0 : r1;
          if (pass == 0 && !r) continue;        // no first filter in this set: F0 is not read
          if (r) {
            const uint32_t s = (uint32_t)dv_sgr_params[set][2 + pass];
            const int n = (2 * r + 1) * (2 * r + 1);
            const uint32_t one_by_n = dv_one_by_x[n - 1];
            // A / B at rows -1..bh (pass 0: odd rows only), columns -1..w
            const int rows = pass == 0 ? (bh + 3) / 2 : bh + 2;
            for (int i = tid; i < rows * (w + 2); i += 256) {
              int ri, aj;
              lr_divmod(i, dv_w2, ri, aj);
              const int ai = pass == 0 ? ri * 2 : ri;             // A row index = i + 1, i = -1, 1, 3, ... in the r = 2 pass
              const uint32_t a = pass == 0 ? sh.R5a[ri * kLrAW + aj] : sh.R3a[ri * kLrAW + aj];
              const uint32_t b = pass == 0 ? sh.R5b[ri * kLrAW + aj] : sh.R3b[ri * kLrAW + aj];
              lr_ab<sizeof(PIX) == 1>(sh.xbx, a, b, bd, s, one_by_n, n, &sh.A[ai * kLrAW + aj], &sh.B[ai * kLrAW + aj]);
            }
          }
          __syncthreads();
          for (int i = tid; i < bh * w; i += 256) {
            int py, px;
            lr_divmod(i, dv_w, py, px);
            const int u_px = sh.S[(py + 3) * kLrSW + px + 3];
            int32_t fv = 0;
            if (r) {
              int a = 0, b = 0;       // a <= 32 * 256, b <= 32 * 2^18: a * u + b < 2^31
              if (pass == 0) {
                if (py & 1) {       // odd row: its own A / B row, weights 6 5 5
                  const uint16_t *Ar = &sh.A[(py + 1) * kLrAW + px + 1]; const int32_t *Br = &sh.B[(py + 1) * kLrAW + px + 1];
                  a = 6 * Ar[0] + 5 * (Ar[-1] + Ar[1]); b = 6 * Br[0] + 5 * (Br[-1] + Br[1]);
                } else {
                  const uint16_t *Au = &sh.A[py * kLrAW + px + 1], *Ad = &sh.A[(py + 2) * kLrAW + px + 1];
                  const int32_t *Bu = &sh.B[py * kLrAW + px + 1], *Bd = &sh.B[(py + 2) * kLrAW + px + 1];
                  a = 6 * (Au[0] + Ad[0]) + 5 * (Au[-1] + Au[1] + Ad[-1] + Ad[1]);
                  b = 6 * (Bu[0] + Bd[0]) + 5 * (Bu[-1] + Bu[1] + Bd[-1] + Bd[1]);
                }
                { const int sh_f = 8 + ((py & 1) ? 4 : 5) - 4; fv = (a * u_px + b + (1 << (sh_f - 1))) >> sh_f; }
              } else {
                for (int dy = -1; dy <= 1; dy++)
                  for (int dx = -1; dx <= 1; dx++) {
                    const int wt = (dx == 0 || dy == 0) ? 4 : 3;
                    a += wt * sh.A[(py + 1 + dy) * kLrAW + px + 1 + dx];
                    b += wt * sh.B[(py + 1 + dy) * kLrAW + px + 1 + dx];
                  }
                fv = (a * u_px + b + (1 << 8)) >> 9;
              }
            }
            if (pass == 0) { sh.F0[py * kLrMaxW + px] = (uint16_t)fv; continue; }
            const int32_t f0 = r0 ? sh.F0[py * kLrMaxW + px] : 0, f1 = fv;
            if (phase == 0) {
              const int u = u_px << 4;
              const int sv = ((int)sh.SRC[py * kLrMaxW + px] << 4) - u;
              const int av = r0 ? f0 - u : 0, bv = r1 ? f1 - u : 0;
              h00 += av * av; h01 += av * bv; h11 += bv * bv; c0 += av * sv; c1 += bv * sv;
              if (ci == 0) ss += sv * sv;       // the same for every set: only the first one is used
            } else {
              const int w0 = sh.best.xqd0, w1 = sh.best.xqd1, w2 = 128 - w0 - w1;
              const long long u = (long long)u_px << 4;
              long long v = w1 * u;
              v += r0 ? (long long)w0 * f0 : w0 * u;
              v += r1 ? (long long)w2 * f1 : w2 * u;
              const long long o = lr_round2(v, 11);
              const int mx = (1 << bd) - 1;
              lr_out[(ys + py) * stride + x0 + px] = (PIX)(o < 0 ? 0 : (o > mx ? mx : o));
            }
          }
          __syncthreads();
        }
        if (phase == 0) {
          const int32_t v[6] = {h00, h01, h11, c0, c1, ss};
#pragma unroll
          for (int k = 0; k < 6; k++) {
            if (k == 5 && ci != 0) continue;
            // exact 64-bit warp sum from two 32-bit hardware reductions: v = (v >> 16) * 65536 + (v & 0xffff)
            const int lo = __reduce_add_sync(0xffffffffu, v[k] & 0xffff), hi = __reduce_add_sync(0xffffffffu, v[k] >> 16);
            const long long tot = (long long)hi * 65536 + lo;
            if ((tid & 31) == 0) sh.acc[tid >> 5][ci][k] += (unsigned long long)tot;
          }
        }
      }
      __syncthreads();
      ys = ye;
    }
    if (phase == 0) {
      __syncthreads();
      if (tid < B200_LR_CANDIDATES_FULL * 6) {      // fold the eight per-warp slots into slot 0
        unsigned long long t = 0;
        for (int wv = 0; wv < 8; wv++) t += (&sh.acc[wv][0][0])[tid];
        (&sh.acc[0][0][0])[tid] = t;
      }
      __syncthreads();
      if (tid == 0) {
        B200LrUnit best = {0, 0, 0, 0};
        const long long S = (long long)sh.acc[0][0][5];
        long long bestJ = ((S >> 8) << 8) + (((long long)f->lambda_q8 * B200_LR_FLAG_COST0) >> 8);
        const long long npx = (long long)w * (y1 - y0);
        for (int ci = 0; ci < n_sets; ci++) {
          const int set = LR_SET_OF(ci);
          const int r0 = dv_sgr_params[set][0], r1 = dv_sgr_params[set][1];
          const long long H00 = (long long)sh.acc[0][ci][0], H01 = (long long)sh.acc[0][ci][1], H11 = (long long)sh.acc[0][ci][2];
          const long long C0 = (long long)sh.acc[0][ci][3], C1 = (long long)sh.acc[0][ci][4];
          const long long h00 = H00 / npx, h01 = H01 / npx, h11 = H11 / npx, c0 = C0 / npx, c1 = C1 / npx;
          long long xq0 = 0, xq1 = 0;
          if (!r0) { if (h11 <= 0) continue; xq1 = lr_round_div(c1 * 128, h11); }
          else if (!r1) { if (h00 <= 0) continue; xq0 = lr_round_div(c0 * 128, h00); }
          else {
            const long long det = h00 * h11 - h01 * h01;
            if (det <= 0) continue;
            xq0 = lr_round_div((c0 * h11 - c1 * h01) * 128, det);
            xq1 = lr_round_div((c1 * h00 - c0 * h01) * 128, det);
          }
          xq0 = xq0 < -512 ? -512 : (xq0 > 512 ? 512 : xq0);
          xq1 = xq1 < -512 ? -512 : (xq1 > 512 ? 512 : xq1);
          int xqd0, xqd1;
          if (!r0) { xqd0 = 0; xqd1 = lr_clamp(128 - (int)xq1, -32, 95); }
          else if (!r1) { xqd0 = lr_clamp((int)xq0, -96, 31); xqd1 = lr_clamp(128 - xqd0, -32, 95); }
          else { xqd0 = lr_clamp((int)xq0, -96, 31); xqd1 = lr_clamp(128 - xqd0 - (int)xq1, -32, 95); }
          const long long e0 = r0 ? xqd0 : 0, e1 = r1 ? 128 - xqd0 - xqd1 : 0;
          const long long E = S * 16384 - 256 * (e0 * C0 + e1 * C1) + e0 * e0 * H00 + 2 * e0 * e1 * H01 + e1 * e1 * H11;
          const long long sse = E <= 0 ? 0 : (E >> 22);
          int bits = 4;
          if (r0) bits += lr_subexp_bits(-96, 32, -32, xqd0);
          if (r1) bits += lr_subexp_bits(-32, 96, 31, xqd1);
          const long long J = (sse << 8) + (((long long)f->lambda_q8 * (B200_LR_FLAG_COST1 + 256 * bits)) >> 8);
          if (J < bestJ) { bestJ = J; best.use = 1; best.set = (int8_t)set; best.xqd0 = (int8_t)xqd0; best.xqd1 = (int8_t)xqd1; }
        }
        sh.best = best;
        job.lr[plane][ur * nuc + uc] = best;
      }
      __syncthreads();
    }
  }
}

// ---- fused deblocking + CDEF: one CTA (256 threads) per 64x64 superblock, all planes; grid.y = job -------------------
// The reconstruction of the superblock and a 7-sample halo is staged ONCE in shared memory; both deblocking passes
// (vertical edges, then horizontal edges: AV1 spec 7.14) run on the staged copy - redundantly on the halo, so that the
// superblock plus the 2-sample border CDEF reads hold the final deblocked samples without any exchange between CTAs -
// then the CDEF direction search, the strength search against the source and the winning filter run on the same copy
// (spec 7.15).  The frame is read once (reconstruction, source) and written once (CDEF output); of the deblocked frame
// only the four rows per superblock row that loop restoration reads across stripe boundaries (spec 7.17.6) are kept
// (job.dbrows).  Why 7: an edge filter modifies at most 6 samples on either side and reads 7; edges of one pass never
// overlap (the filter length is bounded by the smaller transform on either side), so a sample of the inner region
// depends on the edges within 6 samples of it and on 7 samples beyond those edges at most.
// Phase 1: stage + deblock.  Phase 2: direction / variance of the 64 8x8 blocks.  Phase 3: squared error of every
// strength index.  Phase 4: apply the winner into job.fin (if present).

template <class PIX>
__global__ void __launch_bounds__(256, 4) b200_dbcdef_kernel(const B200RdoJob *jobs, const unsigned char *tmaps /* CUtensorMap[job][3] of the rec planes, or null */,
                                                          int tma_x_shift /* 1: the maps describe 8-bit planes as 16-bit elements (x coordinate halves) */) {
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
  const int nsb = f->sb_rows * f->sb_cols;
  if ((int)blockIdx.x >= nsb) return;
  const int sbr = blockIdx.x / f->sb_cols, sbc = blockIdx.x % f->sb_cols;
  const int stride = job.stride, W = f->mi_cols * 4, H = f->mi_rows * 4, bd = f->bit_depth, cshift = bd - 8;
  const int tid = threadIdx.x;
  const int x0 = sbc * 64, y0 = sbr * 64;
  __shared__ int s_dir[64], s_var[64], s_skip[64], s_any;
  __shared__ unsigned long long s_sse[8];
  __shared__ int s_best;
  __shared__ uint32_t s_blk_s[64][2], s_blk_d[64][8][3];     // Psychovisual strength search: per 8x8 luma block sum / sum of squares of the source; per index sum, sum of squares of the output and the cross sum
#ifdef B200_EMU
  static uint16_t s_px[3][kFW * kCdefTW];         // staged samples: rows y0-7 .. y0+70, columns x0-7 .. x0+70 (kCdefNA outside the frame)
#else
  // dynamic shared memory: per plane one 128-byte aligned region that holds the working copy (uint16, 78 rows x 80) and, at
  // its end, the TMA landing zone (PIX, box width x 78): the box is expanded IN PLACE (every thread reads its samples of the
  // plane into registers, barrier, writes the uint16 copy), so staging costs no shared memory of its own - 46.5 KB per CTA
  // instead of 69 (8 bit) / 84 KB (10 bit), i.e. four resident CTAs per SM instead of three / two.
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  constexpr int kTmaBoxW = TmaBox<PIX>::kW, kTmaOff = TmaBox<PIX>::kOff;
  constexpr int kLand = ((kTmaBoxW * kTmaBoxH * (int)sizeof(PIX) + 127) / 128) * 128;
  constexpr int kLzOff = kDbPlaneBytes - kLand;        // landing zone at the end of the plane's region (128-byte aligned)
  static_assert(kLzOff >= 0 && kLzOff % 128 == 0, "landing zone must fit the plane region and stay 128-byte aligned");
  uint16_t (*s_px)[kDbPlaneBytes / 2] = reinterpret_cast<uint16_t (*)[kDbPlaneBytes / 2]>(dyn_smem);
  __shared__ __align__(8) unsigned long long s_bar;
#endif
  __shared__ uint8_t s_txs[20 * 20];              // tx_size of the 4x4 units (sbr*16 - 2 .. +17) x (sbc*16 - 2 .. +17)
  __shared__ int16_t s_off[8][2];                 // CDEF tap offsets in the staged region per direction
  __shared__ int16_t s_str[3][8][2];              // per plane and strength index: primary << cshift, secondary (3 -> 4) << cshift
  __shared__ int8_t s_vs[64];                     // per 8x8 block: luma primary-strength scale 0 | 4 .. 16 from its variance (spec 7.15.2)
  if (tid < 16) s_off[tid >> 1][tid & 1] = (int16_t)(dv_cdef_dir[tid >> 1][tid & 1][0] * kCdefTW + dv_cdef_dir[tid >> 1][tid & 1][1]);
  if (tid < 24) {
    const int str = tid < 8 ? f->cdef_y_strength[tid & 7] : f->cdef_uv_strength[tid & 7];
    int sec = str & 3;
    if (sec == 3) sec = 4;
    s_str[tid >> 3][tid & 7][0] = (int16_t)((str >> 2) << cshift);
    s_str[tid >> 3][tid & 7][1] = (int16_t)(sec << cshift);
  }
  if (tid == 0) s_any = 0;
  if (tid < 8) s_sse[tid] = 0;
  for (int i = tid; i < 64 * 2; i += 256) (&s_blk_s[0][0])[i] = 0;
  for (int i = tid; i < 64 * 8 * 3; i += 256) (&s_blk_d[0][0][0])[i] = 0;
  bool staged = false;
#ifndef B200_EMU
  if (tmaps) {
    // one thread arms the barrier with the byte count of the boxes and issues one tensor load per plane
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&s_bar, (uint32_t)(f->planes * kTmaBoxW * kTmaBoxH * (int)sizeof(PIX)));
      for (int p = 0; p < f->planes; p++) {
        const unsigned char *tm = tmaps + ((size_t)blockIdx.y * 3 + p) * 128;
        tmap_acquire(tm);
        tma_load_2d(dyn_smem + p * kDbPlaneBytes + kLzOff, tm, (x0 - kTmaOff) >> tma_x_shift, y0 - kFH, &s_bar);
      }
    }
    // everybody waits for the bytes (bounded: a tile that never lands must not hang the GPU: it is counted in
    // b200_tma_timeouts and the region is staged with ordinary loads instead)
    unsigned spins = 0;
    bool landed = true;
    while (!mbar_try_wait(&s_bar, 0)) { if (++spins > (1u << 20)) { landed = false; break; } }
    if (__syncthreads_or(!landed)) {
      if (tid == 0) atomicAdd(&b200_tma_timeouts, 1u);
    } else {
    // landing zone (native width, zeros outside the frame) -> working copy (uint16, kCdefNA outside the frame), in place
    constexpr int kPer = (kFW * kFW + 255) / 256;
    for (int p = 0; p < f->planes; p++) {
      const PIX *Lz = reinterpret_cast<const PIX *>(dyn_smem + p * kDbPlaneBytes + kLzOff);
      uint16_t v[kPer];
#pragma unroll
      for (int k = 0; k < kPer; k++) {
        const int i = tid + 256 * k, ty = i / kFW, tx = i % kFW;
        const int yy = y0 - kFH + ty, xx = x0 - kFH + tx;
        v[k] = (i < kFW * kFW && yy >= 0 && yy < H && xx >= 0 && xx < W) ? (uint16_t)Lz[ty * kTmaBoxW + tx + (kTmaOff - kFH)] : kCdefNA;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kPer; k++) {
        const int i = tid + 256 * k, ty = i / kFW, tx = i % kFW;
        if (i < kFW * kFW) s_px[p][ty * kCdefTW + tx] = v[k];
      }
    }
    staged = true;
    }
  }
#endif
  if (!staged)
    for (int p = 0; p < f->planes; p++) {
      const PIX *R = (const PIX *)job.rec[p];
      for (int i = tid; i < kFW * kFW; i += 256) {
        const int ty = i / kFW, tx = i % kFW;
        const int yy = y0 - kFH + ty, xx = x0 - kFH + tx;
        s_px[p][ty * kCdefTW + tx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (uint16_t)R[(size_t)yy * stride + xx] : kCdefNA;
      }
    }
  for (int i = tid; i < 400; i += 256) {
    const int r = sbr * 16 - 2 + i / 20, c = sbc * 16 - 2 + i % 20;
    s_txs[i] = (r >= 0 && r < f->mi_rows && c >= 0 && c < f->mi_cols) ? job.mi[r * f->mi_cols + c].tx_size : 0;
  }
  __syncthreads();
#define TXS_AT(r, c) s_txs[((r) - (sbr * 16 - 2)) * 20 + ((c) - (sbc * 16 - 2))]
  // ---- deblocking: pass 0 = vertical edges x0 + 4k (k = 0..16) on every staged row, pass 1 = horizontal edges y0 + 4k on
  // the columns x0-2 .. x0+65.  Same edge rules as the frame-level filter of the oracle (orc_deblock_frame).
  for (int pass = 0; pass < 2; pass++) {
    const int span = pass == 0 ? kFW : 68;
    for (int i = tid; i < f->planes * 17 * span; i += 256) {
      int plane, k, j;        // constant divisors (multiply-shift) in either pass
      if (pass == 0) { plane = i / (17 * kFW); k = (i / kFW) % 17; j = i % kFW; }
      else { plane = i / (17 * 68); k = (i / 68) % 17; j = i % 68; }
      const int lvl = plane == 0 ? f->lf_level[pass] : f->lf_level[plane + 1];
      if (!lvl) continue;
      int x, y;
      if (pass == 0) { x = x0 + 4 * k; y = y0 - kFH + j; } else { x = x0 - 2 + j; y = y0 + 4 * k; }
      if (x < 0 || y < 0 || (x & ~3) >= f->width || (y & ~3) >= f->height) continue;     // 4x4 units that start inside the visible frame
      const int r = y >> 2, c = x >> 2;
      if (pass == 0 ? c == 0 : r == 0) continue;
      const int tx = 4 << TXS_AT(r, c);
      if (pass == 0 ? (x & (tx - 1)) : (y & (tx - 1))) continue;
      const int ptx = 4 << (pass == 0 ? TXS_AT(r, c - 1) : TXS_AT(r - 1, c));
      const int base = min(tx, ptx);
      int flen = plane == 0 ? min(base, 16) : min(base, 8);
      if (plane > 0 && flen == 8) flen = 6;
      deblock_sample<uint16_t>(&s_px[plane][(y - (y0 - kFH)) * kCdefTW + (x - (x0 - kFH))], pass == 0 ? 1 : kCdefTW, flen, lvl, f->lf_sharpness, bd);
    }
    __syncthreads();
  }
#undef TXS_AT
  // ---- keep what later stages read of the deblocked frame: rows 54..57 of the superblock (stripe boundaries of the loop
  // restoration), and the whole superblock when a full plane was asked for (test seam)
  for (int p = 0; p < f->planes; p++) {
    PIX *DR = (PIX *)job.dbrows[p], *DF = (PIX *)job.deb_full[p];
    if (DR)
      for (int i = tid; i < 4 * 64; i += 256) {
        const int ly = 54 + (i >> 6), lx = i & 63;
        if (y0 + ly < H && x0 + lx < W) DR[(size_t)db_row_slot(y0 + ly) * stride + x0 + lx] = (PIX)s_px[p][(ly + kFH) * kCdefTW + lx + kFH];
      }
    if (DF)
      for (int i = tid; i < 4096; i += 256) {
        const int ly = i >> 6, lx = i & 63;
        if (y0 + ly < H && x0 + lx < W) DF[(size_t)(y0 + ly) * stride + x0 + lx] = (PIX)s_px[p][(ly + kFH) * kCdefTW + lx + kFH];
      }
  }
  if (tid < 64) {
    const int r = sbr * 16 + (tid >> 3) * 2, c = sbc * 16 + (tid & 7) * 2;
    int skip = 1;
    if (r < f->mi_rows && c < f->mi_cols) {
      const B200Mi *m = &job.mi[r * f->mi_cols + c];
      skip = (m[0].flags & 1) && (m[1].flags & 1) && (m[f->mi_cols].flags & 1) && (m[f->mi_cols + 1].flags & 1);
      if (!skip) {
        cdef_direction(&s_px[0][((tid >> 3) * 8 + kFH) * kCdefTW + (tid & 7) * 8 + kFH], kCdefTW, bd, &s_dir[tid], &s_var[tid]);
        const int var = s_var[tid];
        s_vs[tid] = (int8_t)(var ? 4 + ((var >> 6) ? min(floor_log2_d((unsigned)(var >> 6)), 12) : 0) : 0);
      }
    }
    s_skip[tid] = skip;
  }
  {   // any non-skip 4x4 unit in the superblock (oracle)
    const int r = sbr * 16 + (tid >> 4), c = sbc * 16 + (tid & 15);
    if (r < f->mi_rows && c < f->mi_cols && !(job.mi[r * f->mi_cols + c].flags & 1)) s_any = 1;
  }
  __syncthreads();
  if (!s_any || !f->enable_cdef) {
    if (tid == 0) job.cdef_idx[blockIdx.x] = -1;
    if (job.fin[0])
      for (int p = 0; p < f->planes; p++)
        for (int i = tid; i < 4096; i += 256) {
          const int ly = i >> 6, lx = i & 63;
          if (y0 + ly < H && x0 + lx < W) ((PIX *)job.fin[p])[(size_t)(y0 + ly) * stride + x0 + lx] = (PIX)s_px[p][(ly + kFH) * kCdefTW + lx + kFH];
        }
    return;
  }
  // ---- strength search: every pixel loads its taps once and evaluates all strength indices on them (squared error
  // against the source per index), then the winner is applied.  Skipped 8x8 blocks are not filtered and do not count.
  const int nidx = 1 << f->cdef_bits;              // <= 8
  for (int pass = 0; pass < 2; pass++) {           // 0: measure every index, 1: apply the winner
    const bool apply = pass == 1;
    if (apply && !job.fin[0]) break;
    const int best = apply ? s_best : 0;
    unsigned int sse[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < f->planes; p++) {
      const int damping = f->cdef_damping + cshift - (p > 0);
      const uint16_t *T = s_px[p];
      const PIX *S = (const PIX *)job.src[p];
      // Tune::Psychovisual: luma is scored per 8x8 block with the variance-aware error of the trials (oracle cdef_sb_sse): the
      // sums of a block row segment (8 lanes: same block, same row) are reduced with shuffles and added to the block's counters
      const bool psy = !apply && p == 0 && f->psy;
      for (int i = tid; i < 4096; i += 256) {
        const int ly = i >> 6, lx = i & 63;
        const int yy = y0 + ly, xx = x0 + lx;
        const bool inside = yy < H && xx < W;
        const int b = (ly >> 3) * 8 + (lx >> 3);
        const uint16_t *c = T + (ly + kFH) * kCdefTW + lx + kFH;
        const bool act = inside && !s_skip[b];
        int pv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, psrc = 0;       // psy: filtered value per index and the source sample (0: not counted)
        if (inside && !act && apply) ((PIX *)job.fin[p])[(size_t)yy * stride + xx] = (PIX)*c;
        if (act) {
        CdefTaps taps;
        cdef_load_taps(c, &s_off[0][0], s_dir[b], taps);
        const int vscale = p == 0 ? (int)s_vs[b] : -1;      // luma: primary strength scaled by the block variance (spec 7.15.2)
        const int src = apply ? 0 : (int)S[(size_t)yy * stride + xx];
        int sec_cached = 0, sec_sum = 0;       // secondary tap sum of the last secondary strength seen (uniform over the CTA)
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (k >= nidx || (apply && k != best)) continue;
          const int pri0 = s_str[p][k][0], sec = s_str[p][k][1];
          int v;
          if (pri0 == 0 && sec != 0) v = cdef_pixel_tile(T, lx, ly, 0, sec, 0, damping, cshift);      // direction 0 when the primary strength is zero
          else {
            int pri = pri0;
            if (p == 0 && pri) pri = vscale ? (pri * vscale + 8) >> 4 : 0;
            if (!pri && !sec) v = taps.x;
            else {
              int sum = pri ? cdef_pri_sum(taps, pri, damping, cshift) : 0;
              if (sec) {
                if (sec != sec_cached) { sec_sum = cdef_sec_sum(taps, sec, damping); sec_cached = sec; }
                sum += sec_sum;
              }
              v = cdef_finish(taps, sum, pri && sec);
            }
          }
          if (apply) ((PIX *)job.fin[p])[(size_t)yy * stride + xx] = (PIX)v;
          else if (psy) pv[k] = v;
          else { const int d = v - src; sse[k] += (unsigned)(d * d); }
        }
        psrc = src;
        }
        if (psy) {        // uniform over the CTA: every lane takes part in the shuffles, inactive lanes add zeros
          uint32_t a = (uint32_t)psrc, a2 = a * a;
#pragma unroll
          for (int m = 1; m < 8; m <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, m); a2 += __shfl_xor_sync(0xffffffffu, a2, m); }
          if ((lx & 7) == 0 && act) { atomicAdd(&s_blk_s[b][0], a); atomicAdd(&s_blk_s[b][1], a2); }
#pragma unroll
          for (int k = 0; k < 8; k++) {
            if (k >= nidx) continue;
            uint32_t d = (uint32_t)pv[k], d2 = d * d, sd = d * (uint32_t)psrc;
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
              d += __shfl_xor_sync(0xffffffffu, d, m); d2 += __shfl_xor_sync(0xffffffffu, d2, m); sd += __shfl_xor_sync(0xffffffffu, sd, m);
            }
            if ((lx & 7) == 0 && act) { atomicAdd(&s_blk_d[b][k][0], d); atomicAdd(&s_blk_d[b][k][1], d2); atomicAdd(&s_blk_d[b][k][2], sd); }
          }
        }
      }
    }
    if (!apply && f->psy) {
      // one (block, index) per thread: the block's Psychovisual error from its five sums
      __syncthreads();
      for (int i = tid; i < 64 * nidx; i += 256) {
        const int b = i >> f->cdef_bits, k = i & (nidx - 1);
        if (s_skip[b]) continue;
        const unsigned long long dist = b200_psy_dist8x8(s_blk_s[b][0], s_blk_d[b][k][0], s_blk_s[b][1], s_blk_d[b][k][1], s_blk_d[b][k][2], bd);
        atomicAdd(&s_sse[k], dist);
      }
    }
    if (!apply) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (k >= nidx) continue;
        unsigned long long v = sse[k];
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((tid & 31) == 0) atomicAdd(&s_sse[k], v);
      }
      __syncthreads();
      if (tid == 0) {
        int bi = 0;
        unsigned long long bs = s_sse[0];
        for (int k = 1; k < nidx; k++) if (s_sse[k] < bs) { bs = s_sse[k]; bi = k; }
        s_best = bi;
        job.cdef_idx[blockIdx.x] = (int8_t)bi;
      }
      __syncthreads();
    }
  }
}

// ---- deblocking level search (frame.lf_search = !fast_deblock, ravif SpeedTweaks av1encoder.rs:586) -----------------
// One CTA per SAMPLED superblock (every fourth superblock row and column: 1/16 of the frame); grid.y = job.  Per plane the
// reconstruction of the superblock + halo is staged once, every candidate level (records.h b200_lf_candidate) is applied to a
// working copy with the same two-pass edge code as the fused kernel, and the squared error against the source over the
// superblock goes to job.lf_sse[plane][candidate].  b200_lf_pick_kernel then takes the minimum per plane (ties: the earlier
// candidate) and writes the levels into the job's frame (read by the fused kernel) and into job.lf_out (frame header).
// Oracle: orc_lf_search (oracle/av1_filters.c).
template <class PIX>
__global__ void __launch_bounds__(256) b200_lf_search_kernel(const B200RdoJob *jobs) {
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
  if (!f->lf_search) return;
  const int nsb = f->sb_rows * f->sb_cols;
  if ((int)blockIdx.x >= nsb) return;
  const int sbr = blockIdx.x / f->sb_cols, sbc = blockIdx.x % f->sb_cols;
  if (!b200_lf_sampled(sbr, sbc)) return;
  const int stride = job.stride, W = f->mi_cols * 4, H = f->mi_rows * 4, bd = f->bit_depth;
  const int tid = threadIdx.x;
  const int x0 = sbc * 64, y0 = sbr * 64;
  __shared__ uint16_t s_org[kFW * kCdefTW], s_wrk[kFW * kCdefTW];
  __shared__ uint8_t s_txs[20 * 20];
  for (int i = tid; i < 400; i += 256) {
    const int r = sbr * 16 - 2 + i / 20, c = sbc * 16 - 2 + i % 20;
    s_txs[i] = (r >= 0 && r < f->mi_rows && c >= 0 && c < f->mi_cols) ? job.mi[r * f->mi_cols + c].tx_size : 0;
  }
#define TXS_AT(r, c) s_txs[((r) - (sbr * 16 - 2)) * 20 + ((c) - (sbc * 16 - 2))]
  for (int plane = 0; plane < f->planes; plane++) {
    const int base = plane == 0 ? f->lf_level[0] : f->lf_level[plane + 1];
    if (!base) continue;                   // uniform: nothing to search for this plane
    __syncthreads();
    const PIX *R = (const PIX *)job.rec[plane];
    for (int i = tid; i < kFW * kFW; i += 256) {
      const int ty = i / kFW, tx = i % kFW;
      const int yy = y0 - kFH + ty, xx = x0 - kFH + tx;
      s_org[ty * kCdefTW + tx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (uint16_t)R[(size_t)yy * stride + xx] : kCdefNA;
    }
    __syncthreads();
    const PIX *S = (const PIX *)job.src[plane];
    for (int k = 0; k < B200_LF_CANDIDATES; k++) {
      const int lvl = b200_lf_candidate(base, k);
      for (int i = tid; i < kFW * kCdefTW; i += 256) s_wrk[i] = s_org[i];
      __syncthreads();
      if (lvl)
        for (int pass = 0; pass < 2; pass++) {
          const int span = pass == 0 ? kFW : 68;
          for (int i = tid; i < 17 * span; i += 256) {
            const int kk = pass == 0 ? i / kFW : i / 68, j = pass == 0 ? i % kFW : i % 68;
            int x, y;
            if (pass == 0) { x = x0 + 4 * kk; y = y0 - kFH + j; } else { x = x0 - 2 + j; y = y0 + 4 * kk; }
            if (x < 0 || y < 0 || (x & ~3) >= f->width || (y & ~3) >= f->height) continue;
            const int r = y >> 2, c = x >> 2;
            if (pass == 0 ? c == 0 : r == 0) continue;
            const int tx = 4 << TXS_AT(r, c);
            if (pass == 0 ? (x & (tx - 1)) : (y & (tx - 1))) continue;
            const int ptx = 4 << (pass == 0 ? TXS_AT(r, c - 1) : TXS_AT(r - 1, c));
            const int bs = min(tx, ptx);
            int flen = plane == 0 ? min(bs, 16) : min(bs, 8);
            if (plane > 0 && flen == 8) flen = 6;
            deblock_sample<uint16_t>(&s_wrk[(y - (y0 - kFH)) * kCdefTW + (x - (x0 - kFH))], pass == 0 ? 1 : kCdefTW, flen, lvl, f->lf_sharpness, bd);
          }
          __syncthreads();
        }
      unsigned long long sse = 0;
      for (int i = tid; i < 4096; i += 256) {
        const int ly = i >> 6, lx = i & 63;
        if (y0 + ly >= H || x0 + lx >= W) continue;
        const int d = (int)s_wrk[(ly + kFH) * kCdefTW + lx + kFH] - (int)S[(size_t)(y0 + ly) * stride + x0 + lx];
        sse += (unsigned long long)(d * d);
      }
      for (int m = 16; m > 0; m >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, m);
      if ((tid & 31) == 0 && sse) atomicAdd(&job.lf_sse[plane * B200_LF_CANDIDATES + k], sse);
      __syncthreads();
    }
  }
#undef TXS_AT
}
// one thread per job
__global__ void b200_lf_pick_kernel(B200RdoJob *jobs, int n_jobs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  B200RdoJob &job = jobs[j];
  B200Frame *f = &job.frame;
  if (f->lf_search) {
    int out[4] = {f->lf_level[0], f->lf_level[1], f->lf_level[2], f->lf_level[3]};
    for (int plane = 0; plane < f->planes; plane++) {
      const int base = plane == 0 ? f->lf_level[0] : f->lf_level[plane + 1];
      if (!base) continue;
      int best_lvl = base;
      unsigned long long best = 0;
      for (int k = 0; k < B200_LF_CANDIDATES; k++) {
        const unsigned long long sse = job.lf_sse[plane * B200_LF_CANDIDATES + k];
        if (k == 0 || sse < best) { best = sse; best_lvl = b200_lf_candidate(base, k); }
      }
      if (plane == 0) { out[0] = best_lvl; out[1] = best_lvl; } else out[plane + 1] = best_lvl;
    }
    for (int i = 0; i < 4; i++) f->lf_level[i] = out[i];
  }
  if (job.lf_out) for (int i = 0; i < 4; i++) job.lf_out[i] = f->lf_level[i];
}

#ifndef B200_EMU
namespace {
template <class PIX> size_t dbcdef_smem_bytes() { return 3 * (size_t)kDbPlaneBytes; }
template <class PIX> void dbcdef_configure() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  if (done[dev & 63]) return;
  cudaFuncSetAttribute(b200_dbcdef_kernel<PIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dbcdef_smem_bytes<PIX>());
  cudaFuncSetAttribute(b200_dbcdef_kernel<PIX>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  done[dev & 63] = true;
}
}  // namespace
// d_tmaps: device array of CUtensorMap[n_jobs][3] describing the rec planes with an 80 x 78 box (b200_filters_tmap_box),
// or NULL: the fused kernel then stages its region with ordinary loads.
extern "C" int b200_filters_launch(B200RdoJob *d_jobs, int n_jobs, int max_sb, int with_lr, int with_lf_search, int high_bitdepth, const void *d_tmaps,
                                   int tma_x_shift, void *stream, int *launched) {
  if (launched) *launched = 0;
  if (n_jobs <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g2(max_sb, n_jobs), g3(max_sb, n_jobs, 3);       // restoration units <= superblocks
  if (with_lf_search) {
    if (high_bitdepth) b200_lf_search_kernel<uint16_t><<<g2, 256, 0, st>>>(d_jobs);
    else b200_lf_search_kernel<uint8_t><<<g2, 256, 0, st>>>(d_jobs);
    B200_MARK_F("lf_search", st);
    b200_lf_pick_kernel<<<(n_jobs + 63) / 64, 64, 0, st>>>(d_jobs, n_jobs);
    B200_MARK_F("lf_pick", st);
  }
  if (high_bitdepth) {
    dbcdef_configure<uint16_t>();
    b200_dbcdef_kernel<uint16_t><<<g2, 256, dbcdef_smem_bytes<uint16_t>(), st>>>(d_jobs, (const unsigned char *)d_tmaps, 0);
    B200_MARK_F("deblock_cdef", st);
    if (with_lr) { b200_lr_kernel<uint16_t><<<g3, 256, 0, st>>>(d_jobs); B200_MARK_F("lr_search", st); }
  } else {
    dbcdef_configure<uint8_t>();
    b200_dbcdef_kernel<uint8_t><<<g2, 256, dbcdef_smem_bytes<uint8_t>(), st>>>(d_jobs, (const unsigned char *)d_tmaps, tma_x_shift);
    B200_MARK_F("deblock_cdef", st);
    if (with_lr) { b200_lr_kernel<uint8_t><<<g3, 256, 0, st>>>(d_jobs); B200_MARK_F("lr_search", st); }
  }
  if (launched) *launched = 1 + (with_lr ? 1 : 0) + (with_lf_search ? 2 : 0);
  return (int)cudaGetLastError();
}
extern "C" void b200_filters_tmap_box(int high_bitdepth, int *box_w, int *box_h) { *box_w = high_bitdepth ? TmaBox<uint16_t>::kW : TmaBox<uint8_t>::kW; *box_h = kTmaBoxH; }
extern "C" unsigned b200_filters_tma_timeouts(void) {
  unsigned v = 0;
  if (cudaMemcpyFromSymbol(&v, b200_tma_timeouts, sizeof(v)) != cudaSuccess) cudaGetLastError();
  return v;
}
#else
static const B200RdoJob *g_f_jobs;
static bool emu_hbd() { return g_f_jobs[0].frame.bit_depth > 8; }
static void emu_cdef_entry(void *) { if (emu_hbd()) b200_dbcdef_kernel<uint16_t>(g_f_jobs, nullptr, 0); else b200_dbcdef_kernel<uint8_t>(g_f_jobs, nullptr, 0); }
static void emu_lr_entry(void *) { if (emu_hbd()) b200_lr_kernel<uint16_t>(g_f_jobs); else b200_lr_kernel<uint8_t>(g_f_jobs); }
static void emu_lf_entry(void *) { if (emu_hbd()) b200_lf_search_kernel<uint16_t>(g_f_jobs); else b200_lf_search_kernel<uint8_t>(g_f_jobs); }
static void emu_lf_pick_entry(void *) { b200_lf_pick_kernel(const_cast<B200RdoJob *>(g_f_jobs), 1); }
extern "C" int b200_filters_launch_emu(const B200RdoJob *jobs, int max_sb) {
  g_f_jobs = jobs;
  if (jobs[0].frame.lf_search) {
    emu::launch(emu_lf_entry, nullptr, (unsigned)max_sb, 256);
    emu::launch(emu_lf_pick_entry, nullptr, 1, 64);
  }
  emu::launch(emu_cdef_entry, nullptr, (unsigned)max_sb, 256);
  for (g_f_plane = 0; g_f_plane < jobs[0].frame.planes; g_f_plane++)
    if (jobs[0].frame.lr_type[g_f_plane]) emu::launch(emu_lr_entry, nullptr, (unsigned)max_sb, 256);
  return 0;
}
#endif
