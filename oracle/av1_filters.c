/* ORACLE (test infrastructure).  AV1 in-loop filters restated from the AV1 specification:
 * deblocking (spec 7.14) and CDEF (spec 7.15), frame-wide, straightforward loops.
 *
 * In the reference these run inside rav1e (deblock.rs / cdef.rs [EXT: crates.io rav1e 0.8.1]); ravif
 * only switches them on/off (cdef/lrf flags, /root/reference/ravif/src/av1encoder.rs:589-590; deblocking
 * is always on in rav1e).  The filters are NORMATIVE: the decoder applies the same arithmetic, so they
 * are pinned sample-exactly by dav1d decoding the emitted bitstream (tests/test_filters.py).  The
 * strength/level CHOICES are encoder freedom (parity unpinned against rav1e).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../cavif_rs_b200/csrc/records.h"

static const uint8_t k_tx_px[5] = {4, 8, 16, 32, 64};
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---------------------------------------------------------------------------------------------
 * Deblocking: one edge sample position.  px points at q0; step = distance between taps
 * (1 for vertical edges, stride for horizontal edges).  filter_len: 4, 6 (chroma), 8, 16. */
static void deblock_sample(uint16_t *px, int step, int filter_len, int lvl, int sharp, int bd) {
  const int shift = sharp > 4 ? 2 : (sharp > 0 ? 1 : 0);
  int limit = sharp > 0 ? clip3(1, 9 - sharp, lvl >> shift) : (lvl >> shift < 1 ? 1 : lvl >> shift);
  const int blimit = (2 * (lvl + 2) + limit) << (bd - 8);
  const int thresh = (lvl >> 4) << (bd - 8);
  limit <<= (bd - 8);
  const int one = 1 << (bd - 8);
  int p[7], q[7];
  const int taps = filter_len == 4 ? 2 : (filter_len == 6 ? 3 : (filter_len == 8 ? 4 : 7));
  for (int i = 0; i < taps; i++) { q[i] = px[i * step]; p[i] = px[-(i + 1) * step]; }
  const int hev = iabs(p[1] - p[0]) > thresh || iabs(q[1] - q[0]) > thresh;
  int mask = iabs(p[1] - p[0]) <= limit && iabs(q[1] - q[0]) <= limit && (iabs(p[0] - q[0]) * 2 + iabs(p[1] - q[1]) / 2) <= blimit;
  if (filter_len >= 6) mask = mask && iabs(p[2] - p[1]) <= limit && iabs(q[2] - q[1]) <= limit;
  if (filter_len >= 8) mask = mask && iabs(p[3] - p[2]) <= limit && iabs(q[3] - q[2]) <= limit;
  if (!mask) return;
  int flat = 0, flat2 = 0;
  if (filter_len >= 6) {
    flat = iabs(p[1] - p[0]) <= one && iabs(q[1] - q[0]) <= one && iabs(p[2] - p[0]) <= one && iabs(q[2] - q[0]) <= one;
    if (filter_len >= 8) flat = flat && iabs(p[3] - p[0]) <= one && iabs(q[3] - q[0]) <= one;
  }
  if (filter_len == 16)
    flat2 = iabs(p[4] - p[0]) <= one && iabs(q[4] - q[0]) <= one && iabs(p[5] - p[0]) <= one && iabs(q[5] - q[0]) <= one &&
            iabs(p[6] - p[0]) <= one && iabs(q[6] - q[0]) <= one;
  if (filter_len == 4 || !flat) {
    /* narrow filter */
    const int lo = -(1 << (bd - 1)), hi = (1 << (bd - 1)) - 1, off = 0x80 << (bd - 8);
    const int ps1 = p[1] - off, ps0 = p[0] - off, qs0 = q[0] - off, qs1 = q[1] - off;
    int f = hev ? clip3(lo, hi, ps1 - qs1) : 0;
    f = clip3(lo, hi, f + 3 * (qs0 - ps0));
    const int f1 = clip3(lo, hi, f + 4) >> 3, f2 = clip3(lo, hi, f + 3) >> 3;
    px[0] = (uint16_t)(clip3(lo, hi, qs0 - f1) + off);
    px[-step] = (uint16_t)(clip3(lo, hi, ps0 + f2) + off);
    if (!hev) {
      const int f3 = (f1 + 1) >> 1;
      px[step] = (uint16_t)(clip3(lo, hi, qs1 - f3) + off);
      px[-2 * step] = (uint16_t)(clip3(lo, hi, ps1 + f3) + off);
    }
    return;
  }
  /* wide filters: n modified samples per side; window of 2n+1 taps (weights 2 near the centre) */
  int n, lg;
  if (filter_len == 16 && flat2) { n = 6; lg = 4; }
  else if (filter_len == 6) { n = 2; lg = 3; }
  else { n = 3; lg = 3; }
  int s[14], o[12];
  for (int i = 0; i <= n; i++) { s[n - i] = p[i]; s[n + 1 + i] = q[i]; }   /* s[0] = p_n ... s[n] = p0, s[n+1] = q0 ... s[2n+1] = q_n */
  for (int i = -n; i < n; i++) {   /* output sample index i: -n..-1 -> p, 0..n-1 -> q ; position in s = i + n + 1 */
    int t = 0;
    for (int j = -n; j <= n; j++) {
      const int k = clip3(-(n + 1), n, i + j);
      const int tap = (n == 3) ? (j == 0 ? 2 : 1) : (iabs(j) < 2 ? 2 : 1);
      t += tap * s[k + n + 1];
    }
    o[i + n] = (t + (1 << (lg - 1))) >> lg;
  }
  for (int i = -n; i < n; i++) px[i * step] = (uint16_t)o[i + n];
}

/* Deblock one plane in place with the given levels: pass 0 (vertical edges) then pass 1 (horizontal edges). */
static void deblock_plane(const B200Frame *f, uint16_t *P, const B200Mi *mi, int plane, int lvl_v, int lvl_h) {
  const int stride = f->mi_cols * 4;
  {
    for (int pass = 0; pass < 2; pass++) {
      const int lvl = pass == 0 ? lvl_v : lvl_h;
      if (!lvl) continue;
      for (int r = 0; r < f->mi_rows; r++)
        for (int c = 0; c < f->mi_cols; c++) {
          const int x = c * 4, y = r * 4;
          if (x >= f->width || y >= f->height) continue;          /* onScreen */
          if (pass == 0 ? c == 0 : r == 0) continue;              /* picture edge */
          const int tx = k_tx_px[mi[r * f->mi_cols + c].tx_size];
          if (pass == 0 ? (x & (tx - 1)) : (y & (tx - 1))) continue;   /* not a transform edge */
          const B200Mi *pm = pass == 0 ? &mi[r * f->mi_cols + c - 1] : &mi[(r - 1) * f->mi_cols + c];
          const int ptx = k_tx_px[pm->tx_size];
          const int base = tx < ptx ? tx : ptx;
          int flen = plane == 0 ? (base > 16 ? 16 : base) : (base > 8 ? 8 : base);
          if (plane > 0 && flen == 8) flen = 6;
          for (int i = 0; i < 4; i++) {
            uint16_t *px = pass == 0 ? P + (y + i) * stride + x : P + y * stride + x + i;
            deblock_sample(px, pass == 0 ? 1 : stride, flen, lvl, f->lf_sharpness, f->bit_depth);
          }
        }
    }
  }
}
/* Deblock the whole frame in place with the frame's levels. */
void orc_deblock_frame(const B200Frame *f, uint16_t *const planes[3], const B200Mi *mi) {
  for (int plane = 0; plane < f->planes; plane++)
    deblock_plane(f, planes[plane], mi, plane, plane == 0 ? f->lf_level[0] : f->lf_level[plane + 1], plane == 0 ? f->lf_level[1] : f->lf_level[plane + 1]);
}

/* Deblocking level search (lf_search = !fast_deblock, ravif SpeedTweaks av1encoder.rs:586; rav1e deblock_filter_optimize [EXT]):
 * per plane, the candidate levels around the closed-form level (records.h b200_lf_candidate) are applied to the
 * reconstruction and scored by squared error against the source on the sampled superblocks (b200_lf_sampled, coded area);
 * the lowest error wins, ties keep the earlier candidate.  Luma uses one level for both edge directions.  Updates f->lf_level.
 * The device kernel b200_lf_search_kernel mirrors this. */
void orc_lf_search(B200Frame *f, const uint16_t *const rec[3], const uint16_t *const src[3], const B200Mi *mi) {
  if (!f->lf_search) return;
  const int stride = f->mi_cols * 4, W = stride, H = f->mi_rows * 4;
  uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)W * H);
  int out[4] = {f->lf_level[0], f->lf_level[1], f->lf_level[2], f->lf_level[3]};
  for (int plane = 0; plane < f->planes; plane++) {
    const int base = plane == 0 ? f->lf_level[0] : f->lf_level[plane + 1];
    if (!base) continue;
    int best_lvl = base;
    uint64_t best = 0;
    for (int k = 0; k < B200_LF_CANDIDATES; k++) {
      const int lvl = b200_lf_candidate(base, k);
      memcpy(tmp, rec[plane], sizeof(uint16_t) * (size_t)W * H);
      deblock_plane(f, tmp, mi, plane, lvl, lvl);
      uint64_t sse = 0;
      for (int sbr = 0; sbr < f->sb_rows; sbr++)
        for (int sbc = 0; sbc < f->sb_cols; sbc++) {
          if (!b200_lf_sampled(sbr, sbc)) continue;
          for (int y = sbr * 64; y < sbr * 64 + 64 && y < H; y++)
            for (int x = sbc * 64; x < sbc * 64 + 64 && x < W; x++) {
              const int d = (int)tmp[y * stride + x] - (int)src[plane][y * stride + x];
              sse += (uint64_t)(d * d);
            }
        }
      if (k == 0 || sse < best) { best = sse; best_lvl = lvl; }
    }
    if (plane == 0) { out[0] = best_lvl; out[1] = best_lvl; } else out[plane + 1] = best_lvl;
  }
  free(tmp);
  for (int i = 0; i < 4; i++) f->lf_level[i] = out[i];
}

/* ---------------------------------------------------------------------------------------------
 * CDEF (spec 7.15) */
static const int8_t k_cdef_dir[8][2][2] = {{{-1, 1}, {-2, 2}}, {{0, 1}, {-1, 2}}, {{0, 1}, {0, 2}}, {{0, 1}, {1, 2}},
                                           {{1, 1}, {2, 2}},   {{1, 0}, {2, 1}},  {{1, 0}, {2, 0}}, {{1, 0}, {2, -1}}};
static const int8_t k_pri_taps[2][2] = {{4, 2}, {3, 3}};
static const int8_t k_sec_taps[2] = {2, 1};

static int floor_log2(unsigned v) { int l = 0; while (v >>= 1) l++; return l; }

/* direction search on one 8x8 luma block of the deblocked frame */
void orc_cdef_direction(const uint16_t *img, int stride, int bd, int *dir_out, int *var_out) {
  int cost[8] = {0}, partial[8][15];
  memset(partial, 0, sizeof(partial));
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
  static const int div_table[9] = {0, 840, 420, 280, 210, 168, 140, 120, 105};
  for (int i = 0; i < 8; i++) {
    cost[2] += partial[2][i] * partial[2][i];
    cost[6] += partial[6][i] * partial[6][i];
  }
  cost[2] *= div_table[8];
  cost[6] *= div_table[8];
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

static inline int constrain(int diff, int threshold, int damping) {
  if (!threshold) return 0;
  const int adj = damping - floor_log2((unsigned)threshold) > 0 ? damping - floor_log2((unsigned)threshold) : 0;
  const int mag = iabs(diff);
  const int v = clip3(0, mag, threshold - (mag >> adj));
  return diff < 0 ? -v : v;
}

/* in: deblocked planes, out: CDEF output (both full coded size). cdef_idx per 64x64: -1 = not filtered. */
void orc_cdef_frame(const B200Frame *f, uint16_t *const in[3], uint16_t *const out[3], const B200Mi *mi, const int8_t *cdef_idx) {
  const int stride = f->mi_cols * 4, bd = f->bit_depth, cshift = bd - 8;
  const int W = f->mi_cols * 4, H = f->mi_rows * 4;
  for (int p = 0; p < f->planes; p++) memcpy(out[p], in[p], sizeof(uint16_t) * (size_t)W * H);
  if (!f->enable_cdef) return;
  for (int r = 0; r < f->mi_rows; r += 2)
    for (int c = 0; c < f->mi_cols; c += 2) {
      const int idx = cdef_idx[(r >> 4) * f->sb_cols + (c >> 4)];
      if (idx < 0) continue;
      const B200Mi *m = &mi[r * f->mi_cols + c];
      const int skip = (m[0].flags & 1) && (m[1].flags & 1) && (m[f->mi_cols].flags & 1) && (m[f->mi_cols + 1].flags & 1);
      if (skip) continue;
      int ydir, var;
      orc_cdef_direction(in[0] + r * 4 * stride + c * 4, stride, bd, &ydir, &var);
      for (int p = 0; p < f->planes; p++) {
        const int str = p == 0 ? f->cdef_y_strength[idx] : f->cdef_uv_strength[idx];
        int pri = (str >> 2) << cshift, sec = str & 3;
        if (sec == 3) sec = 4;
        sec <<= cshift;
        const int damping = f->cdef_damping + cshift - (p > 0);
        const int dir = pri == 0 ? 0 : ydir;      /* decided BEFORE the variance adjustment of the luma strength */
        if (p == 0 && pri) {
          const int i = var >> 6 ? (floor_log2((unsigned)(var >> 6)) < 12 ? floor_log2((unsigned)(var >> 6)) : 12) : 0;
          pri = var ? (pri * (4 + i) + 8) >> 4 : 0;
        }
        if (!pri && !sec) continue;
        const uint16_t *I = in[p];
        uint16_t *O = out[p];
        const int8_t *pt = k_pri_taps[(pri >> cshift) & 1];
        for (int i = 0; i < 8; i++)
          for (int j = 0; j < 8; j++) {
            const int y0 = r * 4 + i, x0 = c * 4 + j;
            const int x = I[y0 * stride + x0];
            int sum = 0, mx = x, mn = x;
            for (int k = 0; k < 2; k++)
              for (int sgn = -1; sgn <= 1; sgn += 2) {
                int yy = y0 + sgn * k_cdef_dir[dir][k][0], xx = x0 + sgn * k_cdef_dir[dir][k][1];
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                  const int v = I[yy * stride + xx];
                  sum += pt[k] * constrain(v - x, pri, damping);
                  if (v > mx) mx = v;
                  if (v < mn) mn = v;
                }
                for (int doff = -2; doff <= 2; doff += 4) {
                  const int d2 = (dir + doff) & 7;
                  yy = y0 + sgn * k_cdef_dir[d2][k][0]; xx = x0 + sgn * k_cdef_dir[d2][k][1];
                  if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const int v = I[yy * stride + xx];
                    sum += k_sec_taps[k] * constrain(v - x, sec, damping);
                    if (v > mx) mx = v;
                    if (v < mn) mn = v;
                  }
                }
              }
            O[y0 * stride + x0] = (uint16_t)clip3(mn, mx, x + ((8 + sum - (sum < 0)) >> 4));
          }
      }
    }
}

/* ---------------------------------------------------------------------------------------------
 * CDEF strength search (encoder side, rav1e: rdo_loop_decision [EXT]): for every 64x64 superblock that
 * has at least one non-skip 8x8 block, pick the strength index (0 .. 2^cdef_bits - 1) minimising the
 * error of the CDEF output against the source over all planes (squared error; luma per 8x8 with the Psychovisual metric when
 * frame.psy is set) (ties -> lowest index).
 * Writes cdef_idx[] (-1 where the superblock is entirely skipped) and the filtered planes. */
/* dirvar: optional cache of (direction, variance) per 8x8 block of this superblock (64 entries x 2, dir = -1: not yet
 * computed) - the strength search calls this once per candidate index and the directions do not depend on it */
static uint64_t cdef_sb_sse(const B200Frame *f, uint16_t *const in[3], uint16_t *const src[3], uint16_t *const out[3],
                            const B200Mi *mi, int sbr, int sbc, int idx, int write, int *dirvar) {
  const int stride = f->mi_cols * 4, bd = f->bit_depth, cshift = bd - 8;
  const int W = f->mi_cols * 4, H = f->mi_rows * 4;
  uint64_t sse = 0;
  for (int r = sbr * 16; r < sbr * 16 + 16 && r < f->mi_rows; r += 2)
    for (int c = sbc * 16; c < sbc * 16 + 16 && c < f->mi_cols; c += 2) {
      const B200Mi *m = &mi[r * f->mi_cols + c];
      const int skip = (m[0].flags & 1) && (m[1].flags & 1) && (m[f->mi_cols].flags & 1) && (m[f->mi_cols + 1].flags & 1);
      if (skip) continue;
      int ydir, var;
      int *dv = dirvar ? dirvar + 2 * ((((r >> 1) & 7) << 3) | ((c >> 1) & 7)) : NULL;
      if (dv && dv[0] >= 0) { ydir = dv[0]; var = dv[1]; }
      else {
        orc_cdef_direction(in[0] + r * 4 * stride + c * 4, stride, bd, &ydir, &var);
        if (dv) { dv[0] = ydir; dv[1] = var; }
      }
      for (int p = 0; p < f->planes; p++) {
        const int str = p == 0 ? f->cdef_y_strength[idx] : f->cdef_uv_strength[idx];
        int pri = (str >> 2) << cshift, sec = str & 3;
        if (sec == 3) sec = 4;
        sec <<= cshift;
        const int damping = f->cdef_damping + cshift - (p > 0);
        const int dir = pri == 0 ? 0 : ydir;
        if (p == 0 && pri) {
          const int i = var >> 6 ? (floor_log2((unsigned)(var >> 6)) < 12 ? floor_log2((unsigned)(var >> 6)) : 12) : 0;
          pri = var ? (pri * (4 + i) + 8) >> 4 : 0;
        }
        const uint16_t *I = in[p];
        const int8_t *pt = k_pri_taps[(pri >> cshift) & 1];
        /* Tune::Psychovisual: the luma error of an 8x8 block is the variance-aware error of the trials (rav1e scores the
         * luma plane of the loop-filter decisions with cdef_dist_kernel, rdo_loop_plane_error [EXT]); chroma: squared error */
        const int psy = f->psy && p == 0;
        uint32_t ps = 0, pd = 0, ps2 = 0, pd2 = 0, psd = 0;
        for (int i = 0; i < 8; i++)
          for (int j = 0; j < 8; j++) {
            const int y0 = r * 4 + i, x0 = c * 4 + j;
            const int x = I[y0 * stride + x0];
            int y = x;
            if (pri || sec) {
              int sum = 0, mx = x, mn = x;
              for (int k = 0; k < 2; k++)
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                  int yy = y0 + sgn * k_cdef_dir[dir][k][0], xx = x0 + sgn * k_cdef_dir[dir][k][1];
                  if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const int v = I[yy * stride + xx];
                    sum += pt[k] * constrain(v - x, pri, damping);
                    if (v > mx) mx = v;
                    if (v < mn) mn = v;
                  }
                  for (int doff = -2; doff <= 2; doff += 4) {
                    const int d2 = (dir + doff) & 7;
                    yy = y0 + sgn * k_cdef_dir[d2][k][0]; xx = x0 + sgn * k_cdef_dir[d2][k][1];
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                      const int v = I[yy * stride + xx];
                      sum += k_sec_taps[k] * constrain(v - x, sec, damping);
                      if (v > mx) mx = v;
                      if (v < mn) mn = v;
                    }
                  }
                }
              y = clip3(mn, mx, x + ((8 + sum - (sum < 0)) >> 4));
            }
            const int sv = src[p][y0 * stride + x0];
            if (psy) { ps += (uint32_t)sv; pd += (uint32_t)y; ps2 += (uint32_t)(sv * sv); pd2 += (uint32_t)(y * y); psd += (uint32_t)(sv * y); }
            else { const int d = y - sv; sse += (uint64_t)(d * d); }
            if (write) out[p][y0 * stride + x0] = (uint16_t)y;
          }
        if (psy) sse += b200_psy_dist8x8(ps, pd, ps2, pd2, psd, bd);
      }
    }
  return sse;
}

void orc_cdef_search_frame(const B200Frame *f, uint16_t *const in[3], uint16_t *const src[3], uint16_t *const out[3],
                           const B200Mi *mi, int8_t *cdef_idx) {
  const int W = f->mi_cols * 4, H = f->mi_rows * 4;
  for (int p = 0; p < f->planes; p++) memcpy(out[p], in[p], sizeof(uint16_t) * (size_t)W * H);
  for (int sbr = 0; sbr < f->sb_rows; sbr++)
    for (int sbc = 0; sbc < f->sb_cols; sbc++) {
      int any = 0;
      for (int r = sbr * 16; r < sbr * 16 + 16 && r < f->mi_rows && !any; r++)
        for (int c = sbc * 16; c < sbc * 16 + 16 && c < f->mi_cols; c++)
          if (!(mi[r * f->mi_cols + c].flags & 1)) { any = 1; break; }
      if (!any || !f->enable_cdef) { cdef_idx[sbr * f->sb_cols + sbc] = -1; continue; }
      int best = 0;
      uint64_t bs = UINT64_MAX;
      int dirvar[128];
      for (int k = 0; k < 128; k++) dirvar[k] = -1;
      for (int idx = 0; idx < (1 << f->cdef_bits); idx++) {
        const uint64_t s = cdef_sb_sse(f, in, src, out, mi, sbr, sbc, idx, 0, dirvar);
        if (s < bs) { bs = s; best = idx; }
      }
      cdef_idx[sbr * f->sb_cols + sbc] = (int8_t)best;
      cdef_sb_sse(f, in, src, out, mi, sbr, sbc, best, 1, dirvar);
    }
}
