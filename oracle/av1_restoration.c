/* ORACLE (test infrastructure).  AV1 loop restoration, self-guided projection filter only (AV1 spec 7.17, 7.17.3,
 * 7.17.4, 7.17.6), restated from the specification as straightforward per-pixel loops, plus the encoder-side
 * parameter search this repo uses.
 *
 * In the reference LR runs inside rav1e (lrf.rs [EXT: crates.io rav1e 0.8.1]); ravif only switches it on
 * (`lrf`, /root/reference/ravif/src/av1encoder.rs:589, on for every preset) and picks the sgrproj search
 * complexity (:573).  The FILTER is normative: it is pinned against libaom's av1_selfguided_restoration_c /
 * av1_apply_selfguided_restoration_c (tests/test_restoration.py) and, through the emitted bitstream, sample-exactly
 * by dav1d.  The per-unit parameter CHOICE (set, projection weights) is encoder freedom (parity unpinned vs rav1e).
 *
 * Geometry (restoration unit size 64 = lr_unit_shift 0): unit (ur, uc) covers luma rows [64*ur - 8, 64*ur + 56) and
 * columns [64*uc, 64*uc + 64), the last row / column of units absorbing a remainder of less than half a unit; a
 * unit is filtered stripe by stripe (64-row stripes with the same -8 offset): rows outside the stripe come from the
 * DEBLOCKED frame (at most two of them), rows inside from the CDEF output.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../cavif_rs_b200/csrc/records.h"
#include "../cavif_rs_b200/csrc/av1_tables.h"

typedef struct {
  const uint16_t *deb, *cdef;   /* deblocked (pre-CDEF) and CDEF output planes */
  int stride, w, h, bd;         /* visible plane size */
  int stripe_start, stripe_end;
} lr_t;

/* AV1 spec 7.17.6 get_source_sample */
static inline int lr_sample(const lr_t *c, int x, int y) {
  x = x < 0 ? 0 : (x > c->w - 1 ? c->w - 1 : x);
  y = y < 0 ? 0 : (y > c->h - 1 ? c->h - 1 : y);
  if (y < c->stripe_start) { if (y < c->stripe_start - 2) y = c->stripe_start - 2; return c->deb[y * c->stride + x]; }
  if (y > c->stripe_end) { if (y > c->stripe_end + 2) y = c->stripe_end + 2; return c->deb[y * c->stride + x]; }
  return c->cdef[y * c->stride + x];
}
static inline int64_t round2s(int64_t v, int n) { return n == 0 ? v : ((v + ((int64_t)1 << (n - 1))) >> n); }

/* Raw box sums of the region [x0, x0+w) x [y0, y0+h) (inside ONE stripe): sum and sum of squares over the (2r+1)^2
 * window around every A / B position (-1 .. h) x (-1 .. w), r = 1 at every row, r = 2 at the odd rows only (the r = 2
 * pass of spec 7.17.4 reads no other row).  They do not depend on the parameter set, so the search computes them once. */
typedef struct { uint32_t *a3, *b3, *a5, *b5; int aw; } sgr_sums_t;
static void sgr_box_sums(const lr_t *c, int x0, int y0, int w, int h, sgr_sums_t *s) {
  const int aw = w + 2, sw = w + 6, sh = h + 6;
  s->aw = aw;
  s->a3 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)aw * (h + 2)); s->b3 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)aw * (h + 2));
  s->a5 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)aw * (h + 2)); s->b5 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)aw * (h + 2));
  uint16_t *S = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)sw * sh);       /* samples with the 3-sample border */
  for (int i = 0; i < sh; i++)
    for (int j = 0; j < sw; j++) S[i * sw + j] = (uint16_t)lr_sample(c, x0 + j - 3, y0 + i - 3);
  for (int i = -1; i <= h; i++)
    for (int j = -1; j <= w; j++) {
      const uint16_t *ctr = S + (i + 3) * sw + j + 3;
      uint32_t a = 0, b = 0;
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) { const uint32_t v = ctr[dy * sw + dx]; a += v * v; b += v; }
      s->a3[(i + 1) * aw + j + 1] = a; s->b3[(i + 1) * aw + j + 1] = b;
      if (i & 1) {
        for (int dy = -2; dy <= 2; dy++)
          for (int dx = -2; dx <= 2; dx++) {
            if (dy >= -1 && dy <= 1 && dx >= -1 && dx <= 1) continue;
            const uint32_t v = ctr[dy * sw + dx]; a += v * v; b += v;
          }
        s->a5[(i + 1) * aw + j + 1] = a; s->b5[(i + 1) * aw + j + 1] = b;
      }
    }
  free(S);
}
static void sgr_sums_free(sgr_sums_t *s) { free(s->a3); free(s->b3); free(s->a5); free(s->b5); }

/* box filter process (spec 7.17.4) of both passes for one parameter set from the raw sums.  flt0 / flt1: w*h outputs
 * (left untouched when the set's radius is 0). */
static void sgr_from_sums(const lr_t *c, const sgr_sums_t *bs, int x0, int y0, int w, int h, int set, int32_t *flt0, int32_t *flt1) {
  const int aw = bs->aw;
  int32_t *A = (int32_t *)malloc(sizeof(int32_t) * (size_t)aw * (h + 2));
  int32_t *B = (int32_t *)malloc(sizeof(int32_t) * (size_t)aw * (h + 2));
  for (int pass = 0; pass < 2; pass++) {
    const int r = av1_sgr_params_tbl[set][pass];
    if (r == 0) continue;
    const uint32_t s = (uint32_t)av1_sgr_params_tbl[set][2 + pass];
    const int n = (2 * r + 1) * (2 * r + 1);
    const uint32_t one_by_n = (uint32_t)av1_one_by_x_tbl[n - 1];
    for (int i = -1; i <= h; i++) {
      if (pass == 0 && !(i & 1)) continue;                 /* the r = 2 pass only reads the odd rows of A / B */
      for (int j = -1; j <= w; j++) {
        const int k = (i + 1) * aw + j + 1;
        const uint32_t a = pass == 0 ? bs->a5[k] : bs->a3[k], b = pass == 0 ? bs->b5[k] : bs->b3[k];
        const uint32_t a_s = (uint32_t)round2s(a, 2 * (c->bd - 8)), d = (uint32_t)round2s(b, c->bd - 8);
        const int64_t pp = (int64_t)a_s * n - (int64_t)d * d;
        const uint64_t p = pp < 0 ? 0 : (uint64_t)pp;
        const uint64_t z = (p * s + (1u << 19)) >> 20;
        const uint32_t a2 = (uint32_t)av1_x_by_xplus1_tbl[z > 255 ? 255 : z];
        A[k] = (int32_t)a2;
        B[k] = (int32_t)round2s((int64_t)(256 - a2) * b * one_by_n, 12);
      }
    }
    int32_t *F = pass == 0 ? flt0 : flt1;
    for (int i = 0; i < h; i++) {
      const int shift = (pass == 0 && (i & 1)) ? 4 : 5;
      for (int j = 0; j < w; j++) {
        int64_t a = 0, b = 0;
        for (int dy = -1; dy <= 1; dy++)
          for (int dx = -1; dx <= 1; dx++) {
            int wt;
            if (pass == 0) wt = ((i + dy) & 1) ? (dx == 0 ? 6 : 5) : 0;
            else wt = (dx == 0 || dy == 0) ? 4 : 3;
            if (!wt) continue;
            a += wt * (int64_t)A[(i + dy + 1) * aw + j + dx + 1];
            b += wt * (int64_t)B[(i + dy + 1) * aw + j + dx + 1];
          }
        const int64_t v = a * c->cdef[(y0 + i) * c->stride + x0 + j] + b;
        F[i * w + j] = (int32_t)round2s(v, 8 + shift - 4);
      }
    }
  }
  free(A); free(B);
}
void orc_sgr_region(const lr_t *c, int x0, int y0, int w, int h, int set, int32_t *flt0, int32_t *flt1) {
  sgr_sums_t bs;
  sgr_box_sums(c, x0, y0, w, h, &bs);
  sgr_from_sums(c, &bs, x0, y0, w, h, set, flt0, flt1);
  sgr_sums_free(&bs);
}

/* test hook: filter one region of a plane whose rows all belong to one stripe */
void orc_sgr_region_plane(const uint16_t *deb, const uint16_t *cdef, int stride, int w, int h, int bd, int stripe_start,
                          int stripe_end, int x0, int y0, int rw, int rh, int set, int32_t *flt0, int32_t *flt1) {
  lr_t c = {deb, cdef, stride, w, h, bd, stripe_start, stripe_end};
  orc_sgr_region(&c, x0, y0, rw, rh, set, flt0, flt1);
}

/* spec 7.17.3: final projection of one pixel */
static inline int sgr_project(int u_px, int32_t f0, int32_t f1, int set, int w0, int w1, int bd) {
  const int w2 = 128 - w0 - w1;
  const int64_t u = (int64_t)u_px << 4;
  int64_t v = w1 * u;
  v += av1_sgr_params_tbl[set][0] ? (int64_t)w0 * f0 : w0 * u;
  v += av1_sgr_params_tbl[set][1] ? (int64_t)w2 * f1 : w2 * u;
  const int64_t s = round2s(v, 11);
  const int mx = (1 << bd) - 1;
  return (int)(s < 0 ? 0 : (s > mx ? mx : s));
}

/* rows [y0, y1) of unit row ur / columns [x0, x1) of unit column uc */
static void unit_rect(const B200Frame *f, int ur, int uc, int *x0, int *x1, int *y0, int *y1) {
  const int nur = b200_lr_units(f->height), nuc = b200_lr_units(f->width);
  *y0 = ur * 64 - 8 < 0 ? 0 : ur * 64 - 8;
  *y1 = ur == nur - 1 ? f->height : ur * 64 + 56;
  *x0 = uc * 64;
  *x1 = uc == nuc - 1 ? f->width : uc * 64 + 64;
}

/* bits of decode_signed_subexp_with_ref_bool (spec 4.10.10 / 5.11.58), k = 4, for value v in [lo, hi) against ref */
static int subexp_bits(int lo, int hi, int ref, int v) {
  const int mx = hi - lo, r = ref - lo;
  int x = v - lo;
  int t;   /* recentered value */
  if ((r << 1) <= mx) { t = x > 2 * r ? x : (x >= r ? (x - r) << 1 : ((r - x) << 1) - 1); }
  else { const int r2 = mx - 1 - r, x2 = mx - 1 - x; t = x2 > 2 * r2 ? x2 : (x2 >= r2 ? (x2 - r2) << 1 : ((r2 - x2) << 1) - 1); }
  int i = 0, mk = 0, bits = 0;
  const int k = 4;
  for (;;) {
    const int b2 = i ? k + i - 1 : k, a = 1 << b2;
    if (mx <= mk + 3 * a) {
      const int nsym = mx - mk;          /* NS(nsym) */
      int w = 0; while ((1 << w) <= nsym) w++;        /* w = floor(log2(nsym)) + 1 */
      const int m = (1 << w) - nsym;
      return bits + ((t - mk) < m ? w - 1 : w);
    }
    bits++;
    if (t >= mk + a) { i++; mk += a; } else return bits + b2;
  }
}

static inline int64_t round_div(int64_t a, int64_t b) { return a >= 0 ? (a + b / 2) / b : -((-a + b / 2) / b); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

const int8_t orc_lr_candidate_sets[B200_LR_CANDIDATES] = {B200_LR_SETS};

/* Encoder search for one plane (the device kernel b200_lr_search_kernel mirrors this): per unit, for every candidate
 * set, one pass of the box filters gives the normal equations of the projection; the quantised solution is scored by
 * the quadratic form (no second filter pass), J = SSE * 256 + lambda * bits, against "no restoration". */
void orc_lr_search_plane(const B200Frame *f, int plane, const uint16_t *src, const uint16_t *deb, const uint16_t *cdef,
                         B200LrUnit *units) {
  const int stride = f->mi_cols * 4, nur = b200_lr_units(f->height), nuc = b200_lr_units(f->width);
  (void)plane;
  for (int ur = 0; ur < nur; ur++)
    for (int uc = 0; uc < nuc; uc++) {
      int x0, x1, y0, y1;
      unit_rect(f, ur, uc, &x0, &x1, &y0, &y1);
      const int w = x1 - x0;
      B200LrUnit best = {0, 0, 0, 0};
      int64_t bestJ = 0;
      int have = 0;
      /* stripes of this unit and their raw box sums, shared by every candidate set */
      int n_reg = 0, reg_y[4], reg_h[4];
      lr_t reg_c[4];
      sgr_sums_t reg_s[4];
      for (int ys = y0; ys < y1;) {
        const int stripe = (ys + 8) >> 6;
        lr_t c = {deb, cdef, stride, f->width, f->height, f->bit_depth, stripe * 64 - 8, stripe * 64 + 55};
        const int ye = c.stripe_end + 1 < y1 ? c.stripe_end + 1 : y1;
        reg_c[n_reg] = c; reg_y[n_reg] = ys; reg_h[n_reg] = ye - ys;
        sgr_box_sums(&c, x0, ys, w, ye - ys, &reg_s[n_reg]);
        n_reg++;
        ys = ye;
      }
      /* sgr_complexity_full (ravif SpeedTweaks, av1encoder.rs:573): every one of the 16 parameter sets, else four */
      const int ncand = f->lr_full ? B200_LR_CANDIDATES_FULL : B200_LR_CANDIDATES;
      for (int ci = 0; ci < ncand; ci++) {
        const int set = f->lr_full ? ci : orc_lr_candidate_sets[ci];
        const int r0 = av1_sgr_params_tbl[set][0], r1 = av1_sgr_params_tbl[set][1];
        int64_t H00 = 0, H01 = 0, H11 = 0, C0 = 0, C1 = 0, S = 0;
        for (int rg = 0; rg < n_reg; rg++) {
          const int ys = reg_y[rg], h = reg_h[rg];
          int32_t *f0 = (int32_t *)calloc((size_t)w * h, sizeof(int32_t)), *f1 = (int32_t *)calloc((size_t)w * h, sizeof(int32_t));
          sgr_from_sums(&reg_c[rg], &reg_s[rg], x0, ys, w, h, set, f0, f1);
          for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
              const int64_t u = (int64_t)cdef[(ys + i) * stride + x0 + j] << 4;
              const int64_t s = ((int64_t)src[(ys + i) * stride + x0 + j] << 4) - u;
              const int64_t a = r0 ? f0[i * w + j] - u : 0, b = r1 ? f1[i * w + j] - u : 0;
              H00 += a * a; H01 += a * b; H11 += b * b; C0 += a * s; C1 += b * s; S += s * s;
            }
          free(f0); free(f1);
        }
        if (!have) {     /* "no restoration": distortion of the CDEF output, cost of the use_sgrproj = 0 flag */
          bestJ = ((S >> 8) << 8) + (((int64_t)f->lambda_q8 * B200_LR_FLAG_COST0) >> 8);
          have = 1;
        }
        const int64_t npx = (int64_t)w * (y1 - y0);
        const int64_t h00 = H00 / npx, h01 = H01 / npx, h11 = H11 / npx, c0 = C0 / npx, c1 = C1 / npx;
        int64_t xq0 = 0, xq1 = 0;
        if (!r0) { if (h11 <= 0) continue; xq1 = round_div(c1 * 128, h11); }
        else if (!r1) { if (h00 <= 0) continue; xq0 = round_div(c0 * 128, h00); }
        else {
          const int64_t det = h00 * h11 - h01 * h01;
          if (det <= 0) continue;
          xq0 = round_div((c0 * h11 - c1 * h01) * 128, det);
          xq1 = round_div((c1 * h00 - c0 * h01) * 128, det);
        }
        xq0 = xq0 < -512 ? -512 : (xq0 > 512 ? 512 : xq0);
        xq1 = xq1 < -512 ? -512 : (xq1 > 512 ? 512 : xq1);
        int xqd0, xqd1;
        if (!r0) { xqd0 = 0; xqd1 = clampi(128 - (int)xq1, -32, 95); }
        else if (!r1) { xqd0 = clampi((int)xq0, -96, 31); xqd1 = clampi(128 - xqd0, -32, 95); }
        else { xqd0 = clampi((int)xq0, -96, 31); xqd1 = clampi(128 - xqd0 - (int)xq1, -32, 95); }
        const int64_t e0 = r0 ? xqd0 : 0, e1 = r1 ? 128 - xqd0 - xqd1 : 0;       /* effective weights of flt0 / flt1 */
        /* sum (s - (e0 a + e1 b) / 128)^2, scaled by 128^2 */
        const int64_t E = S * 16384 - 256 * (e0 * C0 + e1 * C1) + e0 * e0 * H00 + 2 * e0 * e1 * H01 + e1 * e1 * H11;
        const int64_t sse = E <= 0 ? 0 : (E >> 22);            /* / 128^2 / 16^2 */
        int bits = 4;
        if (r0) bits += subexp_bits(-96, 32, -32, xqd0);
        if (r1) bits += subexp_bits(-32, 96, 31, xqd1);
        const int64_t J = (sse << 8) + (((int64_t)f->lambda_q8 * (B200_LR_FLAG_COST1 + 256 * bits)) >> 8);
        if (J < bestJ) { bestJ = J; best.use = 1; best.set = (int8_t)set; best.xqd0 = (int8_t)xqd0; best.xqd1 = (int8_t)xqd1; }
      }
      for (int rg = 0; rg < n_reg; rg++) sgr_sums_free(&reg_s[rg]);
      units[ur * nuc + uc] = best;
    }
}

/* Apply the chosen parameters: out = restored plane (visible area; the padding keeps the CDEF samples). */
void orc_lr_apply_plane(const B200Frame *f, const uint16_t *deb, const uint16_t *cdef, const B200LrUnit *units, uint16_t *out) {
  const int stride = f->mi_cols * 4, nur = b200_lr_units(f->height), nuc = b200_lr_units(f->width);
  memcpy(out, cdef, sizeof(uint16_t) * (size_t)stride * f->mi_rows * 4);
  for (int ur = 0; ur < nur; ur++)
    for (int uc = 0; uc < nuc; uc++) {
      const B200LrUnit u = units[ur * nuc + uc];
      if (!u.use) continue;
      int x0, x1, y0, y1;
      unit_rect(f, ur, uc, &x0, &x1, &y0, &y1);
      const int w = x1 - x0;
      for (int ys = y0; ys < y1;) {
        const int stripe = (ys + 8) >> 6;
        lr_t c = {deb, cdef, stride, f->width, f->height, f->bit_depth, stripe * 64 - 8, stripe * 64 + 55};
        const int ye = c.stripe_end + 1 < y1 ? c.stripe_end + 1 : y1, h = ye - ys;
        int32_t *f0 = (int32_t *)calloc((size_t)w * h, sizeof(int32_t)), *f1 = (int32_t *)calloc((size_t)w * h, sizeof(int32_t));
        orc_sgr_region(&c, x0, ys, w, h, u.set, f0, f1);
        for (int i = 0; i < h; i++)
          for (int j = 0; j < w; j++)
            out[(ys + i) * stride + x0 + j] = (uint16_t)sgr_project(cdef[(ys + i) * stride + x0 + j], f0[i * w + j], f1[i * w + j], u.set,
                                                                    u.xqd0, u.xqd1, f->bit_depth);
        free(f0); free(f1);
        ys = ye;
      }
    }
}

void orc_lr_frame(const B200Frame *f, uint16_t *const src[3], uint16_t *const deb[3], uint16_t *const cdef[3],
                  B200LrUnit *const units[3], uint16_t *const out[3]) {
  for (int p = 0; p < f->planes; p++) {
    if (f->lr_type[p]) orc_lr_search_plane(f, p, src[p], deb[p], cdef[p], units[p]);
    if (out && out[p]) {
      if (f->lr_type[p]) orc_lr_apply_plane(f, deb[p], cdef[p], units[p], out[p]);
      else memcpy(out[p], cdef[p], sizeof(uint16_t) * (size_t)f->mi_cols * 4 * f->mi_rows * 4);
    }
  }
}

/* test hook: box filters + projection of one region of a plane that is its own "deblocked" frame (single stripe) */
void orc_sgr_apply_region(const uint16_t *plane, int stride, int w, int h, int bd, int x0, int y0, int rw, int rh, int set,
                          int xqd0, int xqd1, uint16_t *out) {
  lr_t c = {plane, plane, stride, w, h, bd, -64, 1 << 20};
  int32_t *f0 = (int32_t *)calloc((size_t)rw * rh, sizeof(int32_t)), *f1 = (int32_t *)calloc((size_t)rw * rh, sizeof(int32_t));
  orc_sgr_region(&c, x0, y0, rw, rh, set, f0, f1);
  for (int i = 0; i < rh; i++)
    for (int j = 0; j < rw; j++)
      out[i * rw + j] = (uint16_t)sgr_project(plane[(y0 + i) * stride + x0 + j], f0[i * rw + j], f1[i * rw + j], set, xqd0, xqd1, bd);
  free(f0); free(f1);
}
