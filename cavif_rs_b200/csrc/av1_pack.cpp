// Host-side AV1 symbol packer: block records -> AV1 OBUs (sequence header + frame OBU).
//
// In the reference this work happens inside rav1e (ContextWriter / WriterBase range coder /
// header.rs) [EXT: crates.io rav1e 0.8.1, reached only through encode_to_av1,
// /root/reference/ravif/src/av1encoder.rs:749-771].  BASELINE.json north_star keeps entropy
// coding on the CPU: the device produces per-block {mode, tx_type, quantized coeffs} records
// (records.h) and this file turns them into a conformant bitstream.  Written from the AV1
// bitstream specification; conformance is pinned by decoding with dav1d (tests/).
#include "av1_pack.h"

#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#include <atomic>
#include <thread>
#include "av1_tables.h"
#include "av1_tables_extra.h"

namespace {

// ------------------------------------------------------------------------------------------
// Bit writer (headers, MSB first)
struct BitWriter {
  std::vector<uint8_t> buf;
  int bitpos = 0;  // bits used in last byte
  void put(uint32_t v, int n) {
    for (int i = n - 1; i >= 0; i--) {
      if (bitpos == 0) buf.push_back(0);
      buf.back() |= ((v >> i) & 1) << (7 - bitpos);
      bitpos = (bitpos + 1) & 7;
    }
  }
  void align() { bitpos = 0; }
  void trailing() { put(1, 1); bitpos = 0; }
};

void put_leb128(std::vector<uint8_t> &out, uint64_t v) {
  do {
    uint8_t b = v & 0x7f;
    v >>= 7;
    if (v) b |= 0x80;
    out.push_back(b);
  } while (v);
}

// ------------------------------------------------------------------------------------------
// Multi-symbol range encoder (AV1 spec 8.2 / daala entropy coder), 15-bit inverse CDFs.
struct RangeEncoder {
  uint64_t low = 0;
  uint32_t rng = 0x8000;
  int cnt = -9;
  std::vector<uint16_t> pre;  // pre-carry bytes

  void normalize(uint64_t l, uint32_t r) {
    int d = 15 - (31 - __builtin_clz(r));   // 16 - ilog(r)
    int c = cnt;
    int s = c + d;
    if (s >= 0) {
      c += 16;
      uint64_t m = ((uint64_t)1 << c) - 1;
      if (s >= 8) {
        pre.push_back((uint16_t)(l >> c));
        l &= m;
        c -= 8;
        m >>= 8;
      }
      pre.push_back((uint16_t)(l >> c));
      s = c + d - 24;
      l &= m;
    }
    low = l << d;
    rng = r << d;
    cnt = s;
  }
  // icdf: 32768 - cumulative; symbol s of n
  void encode(int s, const uint16_t *icdf, int n) {
    const uint32_t fl = s > 0 ? icdf[s - 1] : 32768;
    const uint32_t fh = icdf[s];
    uint64_t l = low;
    uint32_t r = rng;
    const int N = n - 1;
    if (fl < 32768) {
      uint32_t u = (((r >> 8) * (fl >> 6)) >> 1) + 4 * (N - (s - 1));
      uint32_t v = (((r >> 8) * (fh >> 6)) >> 1) + 4 * (N - s);
      l += r - u;
      r = u - v;
    } else {
      r -= (((r >> 8) * (fh >> 6)) >> 1) + 4 * (N - s);
    }
    normalize(l, r);
  }
  void literal_bit(int bit) {
    static const uint16_t half[2] = {16384, 0};
    encode(bit, half, 2);
  }
  std::vector<uint8_t> finish() {
    uint64_t l = low;
    int c = cnt;
    int s = 10;
    const uint64_t m = 0x3FFF;
    uint64_t e = ((l + m) & ~m) | (m + 1);
    s += c;
    if (s > 0) {
      uint64_t n = ((uint64_t)1 << (c + 16)) - 1;
      do {
        pre.push_back((uint16_t)(e >> (c + 16)));
        e &= n;
        s -= 8;
        c -= 8;
        n >>= 8;
      } while (s > 0);
    }
    std::vector<uint8_t> out(pre.size());
    uint32_t carry = 0;
    for (size_t i = pre.size(); i-- > 0;) {
      carry += pre[i];
      out[i] = (uint8_t)carry;
      carry >>= 8;
    }
    return out;
  }
};

inline void update_cdf(uint16_t *cdf, int val, int n) {
  static const int nsymbs2speed[17] = {0, 0, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};
  const int rate = 3 + (cdf[n] > 15) + (cdf[n] > 31) + nsymbs2speed[n];
  int tmp = 32768;
  for (int i = 0; i < n - 1; i++) {
    tmp = (i == val) ? 0 : tmp;
    if (tmp < cdf[i]) cdf[i] -= (uint16_t)((cdf[i] - tmp) >> rate);
    else cdf[i] += (uint16_t)((tmp - cdf[i]) >> rate);
  }
  cdf[n] += (cdf[n] < 32);
}

// ------------------------------------------------------------------------------------------
struct Cdfs {
  uint16_t kf_y_mode[5][5][14];
  uint16_t uv_mode[2][13][15];
  uint16_t partition[20][11];
  uint16_t skip[3][3];
  uint16_t intra_ext_tx[3][4][13][17];
  uint16_t angle_delta[8][8];
  uint16_t cfl_sign[9];
  uint16_t cfl_alpha[6][17];
  uint16_t txb_skip[5][13][3];
  uint16_t eob_extra[5][2][9][3];
  uint16_t dc_sign[2][3][3];
  uint16_t eob_multi16[2][2][6];
  uint16_t eob_multi32[2][2][7];
  uint16_t eob_multi64[2][2][8];
  uint16_t eob_multi128[2][2][9];
  uint16_t eob_multi256[2][2][10];
  uint16_t eob_multi512[2][2][11];
  uint16_t eob_multi1024[2][2][12];
  uint16_t coeff_base_eob[5][2][4][4];
  uint16_t coeff_base[5][2][42][5];
  uint16_t coeff_br[5][2][21][5];
  uint16_t sgrproj_restore[3];

  void init(int base_q_idx) {
    const int q = base_q_idx <= 20 ? 0 : base_q_idx <= 60 ? 1 : base_q_idx <= 120 ? 2 : 3;
    memcpy(kf_y_mode, av1_default_kf_y_mode_cdf, sizeof(kf_y_mode));
    memcpy(uv_mode, av1_default_uv_mode_cdf, sizeof(uv_mode));
    memcpy(partition, av1_default_partition_cdf, sizeof(partition));
    memcpy(skip, av1_default_skip_cdf, sizeof(skip));
    memcpy(intra_ext_tx, av1_default_intra_ext_tx_cdf, sizeof(intra_ext_tx));
    memcpy(angle_delta, av1_default_angle_delta_cdf, sizeof(angle_delta));
    memcpy(cfl_sign, av1_default_cfl_sign_cdf, sizeof(cfl_sign));
    memcpy(cfl_alpha, av1_default_cfl_alpha_cdf, sizeof(cfl_alpha));
    memcpy(txb_skip, av1_default_txb_skip_cdf[q], sizeof(txb_skip));
    memcpy(eob_extra, av1_default_eob_extra_cdf[q], sizeof(eob_extra));
    memcpy(dc_sign, av1_default_dc_sign_cdf[q], sizeof(dc_sign));
    memcpy(eob_multi16, av1_default_eob_multi16_cdf[q], sizeof(eob_multi16));
    memcpy(eob_multi32, av1_default_eob_multi32_cdf[q], sizeof(eob_multi32));
    memcpy(eob_multi64, av1_default_eob_multi64_cdf[q], sizeof(eob_multi64));
    memcpy(eob_multi128, av1_default_eob_multi128_cdf[q], sizeof(eob_multi128));
    memcpy(eob_multi256, av1_default_eob_multi256_cdf[q], sizeof(eob_multi256));
    memcpy(eob_multi512, av1_default_eob_multi512_cdf[q], sizeof(eob_multi512));
    memcpy(eob_multi1024, av1_default_eob_multi1024_cdf[q], sizeof(eob_multi1024));
    memcpy(coeff_base_eob, av1_default_coeff_base_eob_cdf[q], sizeof(coeff_base_eob));
    memcpy(coeff_base, av1_default_coeff_base_cdf[q], sizeof(coeff_base));
    memcpy(coeff_br, av1_default_coeff_br_cdf[q], sizeof(coeff_br));
    memcpy(sgrproj_restore, av1_default_sgrproj_restore_cdf, sizeof(sgrproj_restore));
  }
};

// block geometry helpers (square + the rectangular sizes of the enum, for completeness)
const uint8_t k_bw4[22] = {1, 1, 2, 2, 2, 4, 4, 4, 8, 8, 8, 16, 16, 16, 32, 32, 1, 4, 2, 8, 4, 16};
const uint8_t k_bh4[22] = {1, 2, 1, 2, 4, 2, 4, 8, 4, 8, 16, 8, 16, 32, 16, 32, 4, 1, 8, 2, 16, 4};
inline int wlog2_4(int bs) { int w = k_bw4[bs], l = 0; while ((1 << l) < w) l++; return l; }
inline int hlog2_4(int bs) { int h = k_bh4[bs], l = 0; while ((1 << l) < h) l++; return l; }
const uint8_t k_sq_bsize[5] = {BLOCK_4X4, BLOCK_8X8, BLOCK_16X16, BLOCK_32X32, BLOCK_64X64};

struct TilePacker {
  const B200Frame &f;
  const B200Records &rec;
  RangeEncoder ec;
  Cdfs cdf;
  int mi_row_start, mi_row_end, mi_col_start, mi_col_end;
  // entropy contexts: above arrays indexed by absolute mi col, left arrays by mi_row & 15
  std::vector<uint8_t> above_level[3], above_dc[3];
  uint8_t left_level[3][16], left_dc[3][16];
  int8_t cdef_written = 0;   // per current 64x64: has cdef_idx been written?
  int ref_sgr_xqd[3][2];     // RefSgrXqd (AV1 spec 5.11.58): previous unit's weights in this tile, per plane

  TilePacker(const B200Frame &fr, const B200Records &r) : f(fr), rec(r) {}

  void sym(int s, uint16_t *c, int n) {
    ec.encode(s, c, n);
    update_cdf(c, s, n);
  }
  const B200Mi &mi(int r, int c) const { return rec.mi[(size_t)r * f.mi_cols + c]; }
  bool avail_u(int r) const { return r > mi_row_start; }
  bool avail_l(int c) const { return c > mi_col_start; }

  void write_partition(int r, int c, int bsize_idx /*index into k_sq_bsize*/, bool split, bool has_rows, bool has_cols) {
    // spec: ctx = left*2 + above with neighbours' block dimensions compared against this level
    const int bsl = bsize_idx;   // Mi_Width_Log2 of the block at this level: 8x8 -> 1
    const int above = avail_u(r) && wlog2_4(mi(r - 1, c).bsize) < bsl;
    const int left = avail_l(c) && hlog2_4(mi(r, c - 1).bsize) < bsl;
    const int ctx = (bsl - 1) * 4 + left * 2 + above;
    uint16_t *pc = cdf.partition[ctx];
    const int nsym = bsize_idx == 1 ? 4 : (bsize_idx == 5 ? 8 : 10);
    if (has_rows && has_cols) {
      sym(split ? 3 : 0, pc, nsym);   // PARTITION_NONE = 0, PARTITION_SPLIT = 3
      return;
    }
    if (!has_rows && !has_cols) return;   // PARTITION_SPLIT inferred
    // one-bit split_or_horz / split_or_vert with a probability gathered from the partition CDF
    auto prob = [&](int s) -> int {   // probability mass of symbol s (15-bit)
      const int hi = s > 0 ? pc[s - 1] : 32768;
      return hi - pc[s];
    };
    int psum = 0;
    // partition symbols: 0 NONE 1 HORZ 2 VERT 3 SPLIT 4 HORZ_A 5 HORZ_B 6 VERT_A 7 VERT_B 8 HORZ_4 9 VERT_4
    // The forced direction is implied; P(split) gathers the mass of every partition type that also
    // splits in the OTHER direction.
    if (has_cols) {   // bottom edge (!has_rows): split_or_horz -> "vertical-alike" mass
      psum = prob(2) + prob(3);
      if (nsym > 4) { psum += prob(4) + prob(6) + prob(7); if (nsym == 10) psum += prob(9); }
    } else {          // right edge (!has_cols): split_or_vert -> "horizontal-alike" mass
      psum = prob(1) + prob(3);
      if (nsym > 4) { psum += prob(4) + prob(5) + prob(6); if (nsym == 10) psum += prob(8); }
    }
    uint16_t tmp[3] = {(uint16_t)(32768 - (32768 - psum)), 0, 0};
    // symbol 1 ("split") has probability psum; icdf[0] = 32768 - P(symbol 0) = psum
    tmp[0] = (uint16_t)psum;
    ec.encode(split ? 1 : 0, tmp, 2);   // no adaptation
  }

  // ---- coefficients -------------------------------------------------------------------------
  void write_golomb(int x) {
    x += 1;
    int len = 32 - __builtin_clz(x);
    for (int i = 0; i < len - 1; i++) ec.literal_bit(0);
    for (int i = len - 1; i >= 0; i--) ec.literal_bit((x >> i) & 1);
  }

  // returns cul_level | dc category << 6 for context update
  void write_coeffs(int plane, int mi_r, int mi_c, int txs /*TX_4X4..TX_32X32*/, int tx_type, int y_mode,
                    int bsize, int *cul_out, int *dc_cat_out) {
    const int n = 4 << txs;           // transform width == height
    const int w4 = n >> 2;
    const int ptype = plane > 0;
    const int txs_ctx = txs;          // square transforms
    const int eob = rec.eob[plane][(size_t)mi_r * f.mi_cols + mi_c];
    // dense layout (oracle): the block's slot range; compact streams (device): the plane's cursor, blocks arrive in
    // coding order
    const int16_t *qc;
    if (rec.stream_base) { qc = stream_cur[plane]; stream_cur[plane] += eob; }
    else qc = rec.coef[plane] + b200_coef_offset(&f, mi_r, mi_c);
    // --- all_zero context
    int ctx;
    {
      int top = 0, left = 0, any_top = 0, any_left = 0, dcs = 0;
      for (int k = 0; k < w4; k++) {
        if (mi_c + k < f.mi_cols) {
          const int a = above_level[plane][mi_c + k];
          top = std::max(top, a);
          any_top |= a | above_dc[plane][mi_c + k];
          const int d = above_dc[plane][mi_c + k];
          dcs += d == 1 ? -1 : d == 2 ? 1 : 0;
        }
        if (mi_r + k < f.mi_rows) {
          const int l = left_level[plane][(mi_r + k) & 15];
          left = std::max(left, l);
          any_left |= l | left_dc[plane][(mi_r + k) & 15];
          const int d = left_dc[plane][(mi_r + k) & 15];
          dcs += d == 1 ? -1 : d == 2 ? 1 : 0;
        }
      }
      const int bw = k_bw4[bsize] * 4, bh = k_bh4[bsize] * 4;
      if (plane == 0) {
        if (bw == n && bh == n) ctx = 0;
        else if (top == 0 && left == 0) ctx = 1;
        else if (top == 0 || left == 0) ctx = 2 + (std::max(top, left) > 3);
        else if (std::max(top, left) <= 3) ctx = 4;
        else if (std::min(top, left) <= 3) ctx = 5;
        else ctx = 6;
      } else {
        ctx = 7 + (any_top != 0) + (any_left != 0);
        if (bw * bh > n * n) ctx += 3;
      }
      *dc_cat_out = dcs;   // temporarily carries dc sign ctx sum
    }
    const int dc_sign_ctx = *dc_cat_out == 0 ? 0 : (*dc_cat_out < 0 ? 1 : 2);
    sym(eob == 0, cdf.txb_skip[txs_ctx][ctx], 2);
    if (eob == 0) { *cul_out = 0; *dc_cat_out = 0; return; }

    if (plane == 0) {
      // transform type (intra): set depends on size & reduced_tx_set; 32x32 -> DCT only (not coded)
      if (txs <= TX_16X16) {
        int set_type, eset;   // set_type: 2 = DTT4_IDTX (5 types), 3 = DTT4_IDTX_1DDCT (7 types)
        if (f.reduced_tx_set || txs == TX_16X16) { set_type = 2; eset = 2; }
        else { set_type = 3; eset = 1; }
        const int nsym = set_type == 2 ? 5 : 7;
        sym(av1_ext_tx_ind_tbl[set_type][tx_type], cdf.intra_ext_tx[eset][txs][y_mode], nsym);
      }
    }
    // --- eob position
    int eob_pt, eob_extra;
    {
      static const int16_t group_start[12] = {0, 1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 513};
      int t = 0;
      while (t < 11 && group_start[t + 1] <= eob) t++;
      eob_pt = t;   // eob 1 -> 1, 2 -> 2, 3..4 -> 3, ...
      eob_extra = eob - group_start[t];
    }
    const int txc = plane == 0 ? b200_tx_class(tx_type) : 0;      // transform class (spec get_tx_class): 0 2-D, 1 HORIZ, 2 VERT
    const int eob_multi_ctx = txc != 0;
    switch (txs) {
      case TX_4X4: sym(eob_pt - 1, cdf.eob_multi16[ptype][eob_multi_ctx], 5); break;
      case TX_8X8: sym(eob_pt - 1, cdf.eob_multi64[ptype][eob_multi_ctx], 7); break;
      case TX_16X16: sym(eob_pt - 1, cdf.eob_multi256[ptype][eob_multi_ctx], 9); break;
      default: sym(eob_pt - 1, cdf.eob_multi1024[ptype][eob_multi_ctx], 11); break;
    }
    {
      static const int8_t offset_bits[12] = {0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9};
      const int nbits = offset_bits[eob_pt];
      if (nbits > 0) {
        int bit = (eob_extra >> (nbits - 1)) & 1;
        sym(bit, cdf.eob_extra[txs_ctx][ptype][eob_pt - 3], 2);
        for (int i = 1; i < nbits; i++) ec.literal_bit((eob_extra >> (nbits - 1 - i)) & 1);
      }
    }
    // --- scatter levels into a padded 2-D array (row-major, stride n + 4, 4 extra rows)
    const uint16_t *scan = txs == TX_4X4 ? av1_default_scan_4x4 : txs == TX_8X8 ? av1_default_scan_8x8
                         : txs == TX_16X16 ? av1_default_scan_16x16 : av1_default_scan_32x32;
    const int stride = n + 4;
    uint8_t levels[(32 + 4) * (32 + 4)];
    memset(levels, 0, (size_t)stride * (n + 4));
    // scan (spec get_scan): default for the 2-D class, row by row (mrow) for TX_CLASS_VERT, column by column (mcol) for HORIZ
    auto scan_pos = [&](int i) -> int { return txc == 0 ? scan[i] : (txc == 2 ? i : (i % n) * n + i / n); };
    for (int i = 0; i < eob; i++) {
      const int pos = scan_pos(i), r = pos / n, c = pos % n;
      const int a = abs(qc[i]);
      levels[r * stride + c] = (uint8_t)std::min(a, 127);
    }
    const int log2n = 2 + txs;
    // --- base levels, reverse scan
    for (int i = eob - 1; i >= 0; i--) {
      const int pos = scan_pos(i), r = pos >> log2n, c = pos & (n - 1);
      const int level = abs(qc[i]);
      const uint8_t *L = levels + r * stride + c;
      if (i == eob - 1) {
        const int c4 = i == 0 ? 0 : (i <= (n * n) / 8 ? 1 : (i <= (n * n) / 4 ? 2 : 3));
        sym(std::min(level, 3) - 1, cdf.coeff_base_eob[txs_ctx][ptype][c4], 3);
      } else {
        // neighbourhood / offset by transform class (spec Sig_Ref_Diff_Offset, Coeff_Base_Ctx_Offset, Coeff_Base_Pos_Ctx_Offset)
        int mag = std::min<int>(L[1], 3) + std::min<int>(L[stride], 3);
        if (txc == 0) mag += std::min<int>(L[stride + 1], 3) + std::min<int>(L[2], 3) + std::min<int>(L[2 * stride], 3);
        else if (txc == 1) mag += std::min<int>(L[2], 3) + std::min<int>(L[3], 3) + std::min<int>(L[4], 3);
        else mag += std::min<int>(L[2 * stride], 3) + std::min<int>(L[3 * stride], 3) + std::min<int>(L[4 * stride], 3);
        int cctx = std::min((mag + 1) >> 1, 4);
        if (txc == 0) {
          if (r == 0 && c == 0) cctx = 0;
          else cctx += (r + c < 2) ? 1 : (r + c < 4) ? 6 : 21;
        } else cctx += 26 + 5 * std::min(txc == 2 ? r : c, 2);
        sym(std::min(level, 3), cdf.coeff_base[txs_ctx][ptype][cctx], 4);
      }
      if (level > 2) {
        int mag = std::min<int>(L[1], 15) + std::min<int>(L[stride], 15) + std::min<int>(txc == 0 ? L[stride + 1] : (txc == 1 ? L[2] : L[2 * stride]), 15);
        mag = std::min((mag + 1) >> 1, 6);
        const bool near = txc == 0 ? (r < 2 && c < 2) : (txc == 1 ? c == 0 : r == 0);
        int bctx = (r == 0 && c == 0) ? mag : (near ? mag + 7 : mag + 14);
        uint16_t *bc = cdf.coeff_br[std::min(txs_ctx, 3)][ptype][bctx];
        const int base_range = level - 1 - 2;
        for (int idx = 0; idx < 12; idx += 3) {
          const int k = std::min(base_range - idx, 3);
          sym(k, bc, 4);
          if (k < 3) break;
        }
      }
    }
    // --- signs + golomb remainder, forward scan
    int cul = 0, dc_cat = 0;
    for (int i = 0; i < eob; i++) {
      const int v = qc[i];
      const int level = abs(v);
      if (!level) continue;
      const int sign = v < 0;
      if (i == 0) { sym(sign, cdf.dc_sign[ptype][dc_sign_ctx], 2); dc_cat = sign ? 1 : 2; }
      else ec.literal_bit(sign);
      if (level > 14) write_golomb(level - 15);
      cul += level;
    }
    *cul_out = std::min(cul, 63);
    *dc_cat_out = dc_cat;
  }

  void set_coef_ctx(int plane, int mi_r, int mi_c, int w4, int h4, int cul, int dc) {
    for (int k = 0; k < w4; k++)
      if (mi_c + k < f.mi_cols) { above_level[plane][mi_c + k] = (uint8_t)cul; above_dc[plane][mi_c + k] = (uint8_t)dc; }
    for (int k = 0; k < h4; k++)
      if (mi_r + k < f.mi_rows) { left_level[plane][(mi_r + k) & 15] = (uint8_t)cul; left_dc[plane][(mi_r + k) & 15] = (uint8_t)dc; }
  }

  void write_block(int r, int c) {
    const B200Mi &m = mi(r, c);
    const int bsize = m.bsize;
    const int bw4 = k_bw4[bsize], bh4 = k_bh4[bsize];
    const int skip = m.flags & 1;
    // skip
    {
      const int a = avail_u(r) ? (mi(r - 1, c).flags & 1) : 0;
      const int l = avail_l(c) ? (mi(r, c - 1).flags & 1) : 0;
      sym(skip, cdf.skip[a + l], 2);
    }
    // cdef index: literal, at the first non-skip block of each 64x64
    if (!skip && f.enable_cdef && !cdef_written) {
      const int idx = rec.cdef_idx ? rec.cdef_idx[(r >> 4) * f.sb_cols + (c >> 4)] : 0;
      for (int i = f.cdef_bits - 1; i >= 0; i--) ec.literal_bit((std::max(idx, 0) >> i) & 1);
      cdef_written = 1;
    }
    // y mode
    {
      const int am = avail_u(r) ? mi(r - 1, c).y_mode : DC_PRED;
      const int lm = avail_l(c) ? mi(r, c - 1).y_mode : DC_PRED;
      sym(m.y_mode, cdf.kf_y_mode[av1_intra_mode_context[am]][av1_intra_mode_context[lm]], 13);
    }
    const bool use_angle = bsize >= BLOCK_8X8;
    if (use_angle && m.y_mode >= V_PRED && m.y_mode <= D67_PRED)
      sym(m.angle & 15, cdf.angle_delta[m.y_mode - V_PRED], 7);
    if (f.planes > 1) {
      const int cfl_allowed = (bw4 <= 8 && bh4 <= 8);   // 4:4:4: block <= 32x32
      sym(m.uv_mode, cdf.uv_mode[cfl_allowed][m.y_mode], cfl_allowed ? 14 : 13);
      if (m.uv_mode == UV_CFL_PRED) {
        const int js = (m.flags >> 1) & 7;
        sym(js, cdf.cfl_sign, 8);
        const int sign_u = (js + 1) / 3, sign_v = (js + 1) % 3;   // 0 zero, 1 neg, 2 pos
        if (sign_u) sym(m.cfl & 15, cdf.cfl_alpha[(sign_u - 1) * 3 + sign_v], 16);
        if (sign_v) sym(m.cfl >> 4, cdf.cfl_alpha[(sign_v - 1) * 3 + sign_u], 16);
      }
      if (use_angle && m.uv_mode >= V_PRED && m.uv_mode <= D67_PRED)
        sym(m.angle >> 4, cdf.angle_delta[m.uv_mode - V_PRED], 7);
    }
    // residual (TX_MODE_LARGEST: one transform per plane for blocks <= 32x32... 64x64 -> 4 x TX_32X32? not emitted)
    if (skip) {
      for (int p = 0; p < f.planes; p++) set_coef_ctx(p, r, c, bw4, bh4, 0, 0);
      return;
    }
    const int txs = m.tx_size;
    const int t4 = 1 << txs;
    for (int p = 0; p < f.planes; p++) {
      for (int y = 0; y < bh4; y += t4)
        for (int x = 0; x < bw4; x += t4) {
          if (r + y >= f.mi_rows || c + x >= f.mi_cols) continue;
          int cul, dc;
          int tx_type = m.tx_type;
          write_coeffs(p, r + y, c + x, txs, tx_type, m.y_mode, bsize, &cul, &dc);
          set_coef_ctx(p, r + y, c + x, t4, t4, cul, dc);
        }
    }
  }

  // ---- loop restoration units (AV1 spec 5.11.57 read_lr / 5.11.58 read_lr_unit), coded at the start of the superblock
  void write_literal(int v, int bits) { for (int i = bits - 1; i >= 0; i--) ec.literal_bit((v >> i) & 1); }
  void write_ns(int n, int v) {                       // NS(n) in tile data
    int w = 0; while ((1 << w) <= n) w++;
    const int m = (1 << w) - n;
    if (v < m) write_literal(v, w - 1);
    else { write_literal((v + m) >> 1, w - 1); write_literal((v + m) & 1, 1); }
  }
  void write_subexp(int num_syms, int k, int v) {     // inverse of decode_subexp_bool
    int i = 0, mk = 0;
    for (;;) {
      const int b2 = i ? k + i - 1 : k, a = 1 << b2;
      if (num_syms <= mk + 3 * a) { write_ns(num_syms - mk, v - mk); return; }
      const int more = v >= mk + a;
      ec.literal_bit(more);
      if (more) { i++; mk += a; } else { write_literal(v - mk, b2); return; }
    }
  }
  static int recenter(int r, int v) { return v > 2 * r ? v : (v >= r ? (v - r) << 1 : ((r - v) << 1) - 1); }
  void write_signed_subexp_with_ref(int lo, int hi, int k, int ref, int v) {   // value range [lo, hi)
    const int mx = hi - lo, r = ref - lo, x = v - lo;
    if ((r << 1) <= mx) write_subexp(mx, k, recenter(r, x));
    else write_subexp(mx, k, recenter(mx - 1 - r, mx - 1 - x));
  }
  void write_lr_units(int r, int c) {
    const int ur = r >> 4, uc = c >> 4;               // unit size 64 = superblock size: at most one unit per superblock
    if (ur >= b200_lr_units(f.height) || uc >= b200_lr_units(f.width)) return;
    for (int p = 0; p < f.planes; p++) {
      if (!f.lr_type[p] || !rec.lr[p]) continue;
      const B200LrUnit u = rec.lr[p][ur * b200_lr_units(f.width) + uc];
      sym(u.use ? 1 : 0, cdf.sgrproj_restore, 2);
      if (!u.use) continue;
      write_literal(u.set, 4);
      const int r0 = av1_sgr_params_tbl[u.set][0], r1 = av1_sgr_params_tbl[u.set][1];
      int v0 = 0, v1;
      if (r0) { v0 = u.xqd0; write_signed_subexp_with_ref(-96, 32, 4, ref_sgr_xqd[p][0], v0); }
      if (r1) { v1 = u.xqd1; write_signed_subexp_with_ref(-32, 96, 4, ref_sgr_xqd[p][1], v1); }
      else v1 = std::min(std::max(128 - v0, -32), 95);
      ref_sgr_xqd[p][0] = v0; ref_sgr_xqd[p][1] = v1;
    }
  }

  void write_sb_tree(int r, int c, int lvl /*4: 64x64 ... 0: 4x4*/) {
    if (r >= f.mi_rows || c >= f.mi_cols) return;
    if (lvl == 0) { write_block(r, c); return; }
    const int n4 = 1 << lvl, hbs = n4 >> 1;
    const bool has_rows = (r + hbs) < f.mi_rows, has_cols = (c + hbs) < f.mi_cols;
    const bool split = mi(r, c).bsize != k_sq_bsize[lvl];
    write_partition(r, c, lvl, split, has_rows, has_cols);
    if (!split) { write_block(r, c); return; }
    write_sb_tree(r, c, lvl - 1);
    write_sb_tree(r, c + hbs, lvl - 1);
    write_sb_tree(r + hbs, c, lvl - 1);
    write_sb_tree(r + hbs, c + hbs, lvl - 1);
  }

  const int16_t *stream_cur[3] = {nullptr, nullptr, nullptr};
  std::vector<uint8_t> run(int tile_row, int tile_col) {
    mi_row_start = f.tile_row_start_sb[tile_row] * 16;
    mi_row_end = std::min(f.tile_row_start_sb[tile_row + 1] * 16, f.mi_rows);
    mi_col_start = f.tile_col_start_sb[tile_col] * 16;
    mi_col_end = std::min(f.tile_col_start_sb[tile_col + 1] * 16, f.mi_cols);
    cdf.init(f.base_q_idx);
    for (int p = 0; p < 3; p++) { ref_sgr_xqd[p][0] = -32; ref_sgr_xqd[p][1] = 31; }     // Sgrproj_Xqd_Mid
    for (int p = 0; p < 3; p++) {
      above_level[p].assign(f.mi_cols + 32, 0);
      above_dc[p].assign(f.mi_cols + 32, 0);
    }
    for (int r = mi_row_start; r < mi_row_end; r += 16) {
      if (rec.stream_base)
        for (int p = 0; p < f.planes; p++) stream_cur[p] = rec.stream_base + rec.stream_off[p][(size_t)(r >> 4) * f.tile_cols + tile_col];
      memset(left_level, 0, sizeof(left_level));
      memset(left_dc, 0, sizeof(left_dc));
      for (int c = mi_col_start; c < mi_col_end; c += 16) {
        cdef_written = 0;
        write_lr_units(r, c);
        write_sb_tree(r, c, 4);
      }
    }
    return ec.finish();
  }
};

int level_for(const B200Frame &f) {
  const int64_t px = (int64_t)f.width * f.height;
  const int tiles = f.tile_cols * f.tile_rows;
  struct L { int idx; int64_t pic; int w, h, tiles; };
  static const L lv[] = {{0, 147456, 2048, 1152, 8},    {1, 278784, 2816, 1584, 8},   {4, 665856, 4352, 2448, 16},
                         {5, 1065024, 5504, 3096, 16},  {8, 2359296, 6144, 3456, 32}, {12, 8912896, 8192, 4352, 64},
                         {16, 35651584, 16384, 8704, 128}};
  for (const L &l : lv)
    if (px <= l.pic && f.width <= l.w && f.height <= l.h && tiles <= l.tiles) return l.idx;
  return 31;
}

void write_sequence_header(const B200Frame &f, BitWriter &b) {
  const int profile = f.planes == 1 ? 0 : 1;   // 4:4:4 -> High profile; monochrome -> Main
  b.put(profile, 3);
  b.put(1, 1);   // still_picture
  b.put(1, 1);   // reduced_still_picture_header
  b.put(level_for(f), 5);
  int wb = 1, hb = 1;
  while ((1 << wb) < f.width) wb++;
  while ((1 << hb) < f.height) hb++;
  b.put(wb - 1, 4);
  b.put(hb - 1, 4);
  b.put(f.width - 1, wb);
  b.put(f.height - 1, hb);
  b.put(0, 1);   // use_128x128_superblock
  b.put(0, 1);   // enable_filter_intra
  b.put(f.enable_intra_edge_filter ? 1 : 0, 1);   // enable_intra_edge_filter
  b.put(0, 1);   // enable_superres
  b.put(f.enable_cdef ? 1 : 0, 1);
  b.put(f.enable_restoration ? 1 : 0, 1);
  // color_config
  b.put(f.bit_depth > 8 ? 1 : 0, 1);               // high_bitdepth
  if (profile != 1) b.put(f.planes == 1 ? 1 : 0, 1);   // mono_chrome
  b.put(f.color_description_present ? 1 : 0, 1);
  if (f.color_description_present) {
    b.put(f.color_primaries, 8);
    b.put(f.transfer_characteristics, 8);
    b.put(f.matrix_coefficients, 8);
  }
  if (f.planes == 1) {
    b.put(f.full_range ? 1 : 0, 1);
  } else if (f.color_description_present && f.color_primaries == 1 && f.transfer_characteristics == 13 &&
             f.matrix_coefficients == 0) {
    // sRGB / identity: full range 4:4:4 implied
  } else {
    b.put(f.full_range ? 1 : 0, 1);
    // profile 1: subsampling 0,0 implied
  }
  if (f.planes > 1) b.put(0, 1);   // separate_uv_delta_q
  b.put(0, 1);   // film_grain_params_present
  b.trailing();
}

int tile_log2(int blk, int target) {
  int k = 0;
  while ((blk << k) < target) k++;
  return k;
}

void write_frame_header(const B200Frame &f, BitWriter &b) {
  b.put(0, 1);   // disable_cdf_update
  b.put(0, 1);   // allow_screen_content_tools
  // frame_size(): no override; superres disabled; render_size:
  b.put(0, 1);   // render_and_frame_size_different
  // tile_info
  {
    const int sb_cols = f.sb_cols, sb_rows = f.sb_rows;
    const int max_w_sb = 4096 >> 6, max_area_sb = (4096 * 2304) >> 12;
    const int min_log2_cols = tile_log2(max_w_sb, sb_cols);
    const int max_log2_cols = tile_log2(1, std::min(sb_cols, 64));
    const int max_log2_rows = tile_log2(1, std::min(sb_rows, 64));
    const int min_log2_tiles = std::max(min_log2_cols, tile_log2(max_area_sb, sb_rows * sb_cols));
    b.put(1, 1);   // uniform_tile_spacing_flag
    for (int k = min_log2_cols; k < max_log2_cols; k++) {
      if (k < f.tile_cols_log2) b.put(1, 1);
      else { b.put(0, 1); break; }
    }
    const int min_log2_rows = std::max(min_log2_tiles - f.tile_cols_log2, 0);
    for (int k = min_log2_rows; k < max_log2_rows; k++) {
      if (k < f.tile_rows_log2) b.put(1, 1);
      else { b.put(0, 1); break; }
    }
    if (f.tile_cols_log2 || f.tile_rows_log2) {
      b.put(0, f.tile_cols_log2 + f.tile_rows_log2);   // context_update_tile_id
      b.put(3, 2);                                      // tile_size_bytes_minus_1
    }
  }
  // quantization_params
  b.put(f.base_q_idx, 8);
  b.put(0, 1);                     // DeltaQYDc delta_coded
  if (f.planes > 1) {
    // separate_uv_delta_q = 0 -> diff_uv_delta = 0
    b.put(0, 1);                   // DeltaQUDc
    b.put(0, 1);                   // DeltaQUAc
  }
  b.put(0, 1);                     // using_qmatrix
  b.put(0, 1);                     // segmentation_enabled
  b.put(0, 1);                     // delta_q_present (base_q_idx > 0)
  // loop_filter_params
  b.put(f.lf_level[0], 6);
  b.put(f.lf_level[1], 6);
  if (f.planes > 1 && (f.lf_level[0] || f.lf_level[1])) {
    b.put(f.lf_level[2], 6);
    b.put(f.lf_level[3], 6);
  }
  b.put(f.lf_sharpness, 3);
  b.put(0, 1);                     // loop_filter_delta_enabled
  // cdef_params
  if (f.enable_cdef) {
    b.put(f.cdef_damping - 3, 2);
    b.put(f.cdef_bits, 2);
    for (int i = 0; i < (1 << f.cdef_bits); i++) {
      b.put(f.cdef_y_strength[i] >> 2, 4);
      b.put(f.cdef_y_strength[i] & 3, 2);
      if (f.planes > 1) {
        b.put(f.cdef_uv_strength[i] >> 2, 4);
        b.put(f.cdef_uv_strength[i] & 3, 2);
      }
    }
  }
  // lr_params
  if (f.enable_restoration) {
    // lr_type: 0 -> RESTORE_NONE, 3 -> RESTORE_SGRPROJ (Remap_Lr_Type); lr_unit_shift = 0: 64x64 units; no lr_uv_shift in 4:4:4
    int uses_lr = 0;
    for (int p = 0; p < f.planes; p++) { b.put(f.lr_type[p] ? 3 : 0, 2); uses_lr |= f.lr_type[p] != 0; }
    if (uses_lr) b.put(0, 1);
  }
  b.put(0, 1);   // tx_mode_select = 0 -> TX_MODE_LARGEST
  b.put(f.reduced_tx_set ? 1 : 0, 1);
}

}  // namespace

static std::atomic<int> g_tile_threads{1};
// host threads b200av1_pack_frame may use for the tiles of ONE frame (default 1)
extern "C" void b200av1_pack_set_tile_threads(int n) { g_tile_threads.store(n < 1 ? 1 : n); }

extern "C" int b200av1_seq_level_idx(const B200Frame *f) { return level_for(*f); }

extern "C" int b200av1_pack_frame(const B200Frame *fp, const B200Records *rec, uint8_t **out, size_t *out_len) {
  const B200Frame &f = *fp;
  if (!out || !out_len || f.base_q_idx < 1 || f.base_q_idx > 255) return 4;
  std::vector<uint8_t> obus;
  // sequence header OBU
  {
    BitWriter b;
    write_sequence_header(f, b);
    obus.push_back((1 << 3) | 2);   // obu_type 1, has_size_field
    put_leb128(obus, b.buf.size());
    obus.insert(obus.end(), b.buf.begin(), b.buf.end());
  }
  // frame OBU
  {
    BitWriter b;
    write_frame_header(f, b);
    b.align();
    std::vector<uint8_t> payload = b.buf;
    const int ntiles = f.tile_cols * f.tile_rows;
    if (ntiles > 1) {
      payload.push_back(0);   // tile_start_and_end_present_flag = 0 + byte alignment
    }
    std::vector<std::vector<uint8_t>> tiles(ntiles);
    // tiles are independent entropy-coding contexts: pack them on g_tile_threads host threads (one large image) or
    // serially (batches, where the caller already runs one packer per frame)
    {
      const int nthreads = std::min(ntiles, std::max(1, g_tile_threads.load()));
      std::atomic<int> next{0};
      auto work = [&]() {
        for (int t; (t = next.fetch_add(1)) < ntiles;) {
          TilePacker tp(f, *rec);
          tiles[t] = tp.run(t / f.tile_cols, t % f.tile_cols);
        }
      };
      if (nthreads <= 1) work();
      else {
        std::vector<std::thread> th;
        for (int k = 0; k < nthreads; k++) th.emplace_back(work);
        for (auto &x : th) x.join();
      }
    }
    for (int t = 0; t < ntiles; t++) {
      if (t != ntiles - 1) {
        const uint32_t sz = (uint32_t)tiles[t].size() - 1;
        for (int k = 0; k < 4; k++) payload.push_back((sz >> (8 * k)) & 0xff);
      }
      payload.insert(payload.end(), tiles[t].begin(), tiles[t].end());
    }
    obus.push_back((6 << 3) | 2);   // OBU_FRAME
    put_leb128(obus, payload.size());
    obus.insert(obus.end(), payload.begin(), payload.end());
  }
  uint8_t *p = (uint8_t *)malloc(obus.size());
  if (!p) return 3;
  memcpy(p, obus.data(), obus.size());
  *out = p;
  *out_len = obus.size();
  return 0;
}
