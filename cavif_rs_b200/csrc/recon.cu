// Stage C of the AV1 intra search: the closed-loop reconstruction chain (the only serial stage).
//
// One CTA per TILE.  Inside a tile the work is split three ways so that the chain is short:
//   * planes: a LUMA warp and a CHROMA warp (U in lane group 0, V in lane group 1) walk the same decided blocks in
//     coding order; chroma only waits for luma where a block uses chroma-from-luma (it needs the reconstructed luma of
//     that block);
//   * superblock rows: ROWS row slots run a wavefront - row R may code superblock C once row R-1 has finished C+1 (the
//     above-right superblock is the furthest reconstruction an intra edge reads);
//   * inside a block the transform runs one column / row per lane (warp-shuffle free: the transpose goes through
//     shared memory), quantisation feeds the inverse row pass from registers.
// Every warp keeps the reconstruction of ITS current superblock in shared memory (64x64 samples per plane, native
// sample width) plus the row above it and the column left of it, so intra edges never touch global memory; a finished
// superblock is written back with 16-byte stores.  The source block and the mode info of the NEXT block are fetched
// while the current one is coded.  Quantised levels leave as a COMPACT stream: per (tile superblock row, plane) the eob
// levels of each transform block in scan order, appended in coding order at the start of that row's slot range of the
// coefficient plane; the host packer consumes them with one cursor per plane (records.h).
//
// Replaces rav1e's encode_block / encode_tx_block chain [EXT: crates.io rav1e 0.8.1] reached through encode_to_av1
// (/root/reference/ravif/src/av1encoder.rs:749-771).  Bit-exact CPU restatement: oracle/av1_encode.c reconstruct_block.
#include "av1_dev.cuh"
#ifndef B200_EMU
#include <mutex>
#endif

namespace {

constexpr int kTS = 64;                 // row stride of a superblock tile in shared memory
#ifdef B200_EMU
#define B200_SPIN() emu::yield()
#else
#define B200_SPIN() __nanosleep(32)
#endif

// Working set of one warp: two 64x64 sample tiles (luma warp: double buffer of its superblock, so that the chroma warp
// can still read the previous one for chroma-from-luma; chroma warp: U and V), their borders, and the per-block scratch.
template <int NMAX, class PIX>
struct ReconWarp {
  alignas(16) PIX T[2][64 * kTS];
  PIX AB[2][136];                      // row above the superblock: index c + 1, c = -1 .. 127
  PIX LC[2][64];                       // column left of the superblock
  alignas(16) PIX srcb[2][NMAX * NMAX];    // source block(s), row-major n x n
  PIX above[2][EOFF + 2 * NMAX + 2], left[2][EOFF + 2 * NMAX + 2];
  PIX ef[2][EOFF + 2 * NMAX + 8];
  PIX pred[2][NMAX * NMAX];
  int32_t buf[NMAX == 32 ? 32 * 33 : 2 * 16 * 17];     // transform transpose staging, N x (N + 1) per lane group
  int16_t qc[2][NMAX * NMAX];          // quantised levels in scan order
  int16_t ac[NMAX * NMAX];             // chroma from luma: (luma << 3) - average
};

// progress counters of one tile (shared memory, touched with atomics + block fences)
template <int ROWS>
struct ReconSync {
  unsigned sb_done[2][2 * ROWS];       // [role][row ring]: superblocks of that row finished AND written back
  unsigned y_blocks[ROWS];             // per slot: luma blocks finished (monotone over the rows of the slot)
  unsigned uv_sbs[ROWS];               // per slot: superblocks finished by the chroma warp (monotone)
};

__device__ __forceinline__ unsigned sync_read(unsigned *p) {
  const unsigned v = atomicAdd(p, 0u);
  return v;
}
// lane 0 polls, the warp follows
__device__ __forceinline__ void wait_ge(unsigned *p, unsigned need) {
  for (;;) {
    unsigned v = lane_id() == 0 ? sync_read(p) : 0u;
    v = __shfl_sync(FULL, v, 0);
    if (v >= need) break;
    B200_SPIN();
  }
  __threadfence_block();
}
__device__ __forceinline__ void publish_add(unsigned *p) {
  __threadfence_block();
  __syncwarp();
  if (lane_id() == 0) atomicAdd(p, 1u);
}

// ---- single-candidate transform path of the reconstruction: lane group g (N lanes) codes one plane ---------------
// The chain is bound by INSTRUCTION FETCH, not by arithmetic (ncu: stall_no_instruction 5.6 warps per issue with every
// 1-D transform inlined at each of its four call sites: 190 KB of code, a different 20 KB slice per block and warp).  Each
// 1-D transform therefore exists ONCE per size, behind a call, with its data in thread-local memory; forward column / row
// passes and inverse row / column passes share those copies, and the hot set of the whole kernel stays cache resident.
template <int N, bool HBD>
__device__ __noinline__ void fwd1d_call(int kind, int32_t *v /* in/out */, int bit) {
  int32_t x[N], o[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = v[i];
  fwd1d<N, HBD>(kind, x, o, bit);
#pragma unroll
  for (int i = 0; i < N; i++) v[i] = o[i];
}
template <int N, bool HBD>
__device__ __noinline__ void inv1d_call(int kind, int32_t *v /* in/out */) {
  int32_t x[N], o[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = v[i];
  inv1d<N, HBD>(kind, x, o);
#pragma unroll
  for (int i = 0; i < N; i++) v[i] = o[i];
}

// src: n x n row-major (stride N); pred: n x n (stride N); rec: destination inside the superblock tile (stride kTS).
// Same arithmetic as trial_batch() in av1_dev.cuh (forward: libaom design; quantiser dead zone 109/256 DC, 98/256 AC;
// inverse: AV1 spec 7.13.3) without the rate model.
template <int N, bool HBD, class PIX>
__device__ __noinline__ int recon_tx(const B200Frame *f, bool active, const PIX *src, const PIX *pred, PIX *rec, int32_t *buf,
                                     int16_t *qc, int tx_type) {
  constexpr int S = N + 1;
  constexpr int shift1 = N == 4 ? 0 : (N == 8 ? 1 : (N == 16 ? 2 : 4));
  constexpr int row_shift = N == 4 ? 0 : (N == 8 ? 1 : 2);
  constexpr int qshift = N == 32 ? 1 : 0;
  const int l = lane_id() % N;
  const int bd = f->bit_depth;
  int vert, horz;
  tx_kernels(tx_type, vert, horz);
  // v[] lives in thread-local memory (it is handed to the out-of-line transforms), so the loops around it stay ROLLED:
  // a rolled loop sits in the L0 instruction cache, sixteen unrolled copies of it do not.
  int32_t v[N];
  // forward column pass
  if (active) {
#pragma unroll 1
    for (int r = 0; r < N; r++) v[r] = ((int)src[r * N + l] - (int)pred[r * N + l]) * 4;
    fwd1d_call<N, HBD>(vert, v, N == 32 ? 12 : 13);
#pragma unroll 1
    for (int r = 0; r < N; r++) buf[r * S + l] = rshift_round_t<HBD>(v[r], shift1);
  }
  __syncwarp();
  // forward row pass, quantiser, dequantiser
  int eob = 0;
  if (active) {
#pragma unroll 1
    for (int c = 0; c < N; c++) v[c] = buf[l * S + c];
    fwd1d_call<N, HBD>(horz, v, (N == 32 || N == 16) ? 12 : 13);
    const uint32_t dc_q = (uint32_t)f->dc_q, ac_q = (uint32_t)f->ac_q;
    const uint32_t ac_rcp = 0xffffffffu / ac_q, ac_off = (ac_q * 98) >> 8;
    const int txc = b200_tx_class(tx_type);      // chroma never uses the 1-D types (tx type derived from the uv mode)
    const int clamp_b = bd + 8;
    uint2 sq = make_uint2(0u, 0u);       // scan indices of four coefficients per table load
#pragma unroll 1
    for (int c = 0; c < N; c++) {
      if ((c & 3) == 0 && txc == 0) sq = scan_quad_load<N>(l, c);
      uint32_t q = ac_q, rcp = ac_rcp, off = ac_off;
      if (c == 0 && l == 0) { q = dc_q; rcp = 0xffffffffu / dc_q; off = (dc_q * 109) >> 8; }
      const int32_t cf = v[c];
      const uint32_t a = ((uint32_t)(cf < 0 ? -cf : cf) << qshift) + off;
      uint32_t lv = __umulhi(a, rcp);
      uint32_t rem = a - lv * q;
      if (rem >= q) { lv++; rem -= q; }
      if (rem >= q) lv++;
      if (lv > 32767) lv = 32767;
      const int32_t dqa = (int32_t)((lv * q) & 0xFFFFFF) >> qshift;
      const int si = txc == 0 ? scan_quad_at<N>(sq, c) : (txc == 2 ? l * N + c : c * N + l);
      qc[si] = (int16_t)(cf < 0 ? -(int)lv : (int)lv);
      v[c] = clamp_bits(cf < 0 ? -dqa : dqa, clamp_b);
      if (lv) eob = max(eob, si + 1);
    }
  }
  eob = group_max<N>(eob);
  // inverse row pass
  if (active && eob) {
    inv1d_call<N, HBD>(horz, v);
#pragma unroll 1
    for (int c = 0; c < N; c++) buf[l * S + c] = rshift_round_t<HBD>(v[c], row_shift);
  }
  __syncwarp();
  if (active) {
    if (eob) {
      const int col_clamp = bd + 6 > 16 ? bd + 6 : 16;
      const int mx = (1 << bd) - 1;
#pragma unroll 1
      for (int r = 0; r < N; r++) v[r] = clamp_bits(buf[r * S + l], col_clamp);
      inv1d_call<N, HBD>(vert, v);
#pragma unroll 1
      for (int r = 0; r < N; r++) {
        int y = (int)pred[r * N + l] + rshift_round_t<HBD>(v[r], 4);
        y = y < 0 ? 0 : (y > mx ? mx : y);
        rec[r * kTS + l] = (PIX)y;
      }
    } else {
#pragma unroll 1
      for (int r = 0; r < N; r++) rec[r * kTS + l] = pred[r * N + l];
    }
  }
  __syncwarp();
  return eob;
}

// ---- intra edges of block (bx, by) [superblock-local samples] from the warp's tile, above row and left column --------
// Same rules as build_edges() (AV1 spec 7.11.2): `avail` from block_avail(); max_lx / max_ly = last coded column / row in
// superblock-local coordinates.
template <class PIX>
__device__ __forceinline__ void edges_from_tile(const PIX *T, const PIX *AB, const PIX *LC, PIX *Aout, PIX *Lout, int bx, int by, int n,
                                                int avail, int bd, int max_lx, int max_ly) {
  const int ha = avail & 1, hl = (avail >> 1) & 1, have_ar = (avail >> 2) & 1, have_bl = (avail >> 3) & 1;
  const int base = 1 << (bd - 1);
  PIX *A = Aout + EOFF, *L = Lout + EOFF;
  // value at (row, col), row >= -1, col >= -1
#define TILE_AT(r, c) ((r) < 0 ? AB[(c) + 1] : ((c) < 0 ? LC[(r)] : T[(r) * kTS + (c)]))
#pragma unroll 1
  for (int i = lane_id(); i < 2 * n; i += 32) {
    PIX a, l;
    if (!ha && hl) a = TILE_AT(by, bx - 1);
    else if (!ha && !hl) a = (PIX)(base - 1);
    else {
      int xx = (i < n || have_ar) ? bx + i : bx + n - 1;
      if (xx > max_lx) xx = max_lx;
      a = TILE_AT(by - 1, xx);
    }
    if (!hl && ha) l = TILE_AT(by - 1, bx);
    else if (!hl && !ha) l = (PIX)(base + 1);
    else {
      int yy = (i < n || have_bl) ? by + i : by + n - 1;
      if (yy > max_ly) yy = max_ly;
      l = TILE_AT(yy, bx - 1);
    }
    A[i] = a; L[i] = l;
  }
  if (lane_id() == 0) {
    PIX tl;
    if (ha && hl) tl = TILE_AT(by - 1, bx - 1);
    else if (ha) tl = TILE_AT(by - 1, bx);
    else if (hl) tl = TILE_AT(by, bx - 1);
    else tl = (PIX)base;
    A[-1] = tl; L[-1] = tl;
  }
#undef TILE_AT
  __syncwarp();
}

template <bool HBD, int NMAX, class PIX>
__device__ __forceinline__ int recon_tx_any(int n, const B200Frame *f, bool active, const PIX *src, const PIX *pred, PIX *rec, int32_t *buf,
                                            int16_t *qc, int tx_type) {
  if (n == 4) return recon_tx<4, HBD, PIX>(f, active, src, pred, rec, buf, qc, tx_type);
  if (n == 8) return recon_tx<8, HBD, PIX>(f, active, src, pred, rec, buf, qc, tx_type);
  if (NMAX == 16 || n == 16) return recon_tx<16, HBD, PIX>(f, active, src, pred, rec, buf, qc, tx_type);
  return recon_tx<(NMAX == 32 ? 32 : 16), HBD, PIX>(f, active, src, pred, rec, buf, qc, tx_type);
}

struct BlockInfo { B200Mi m; uint8_t mode_above, mode_left; };

// One superblock row of a tile, coded by one warp (role 0: luma, role 1: U + V).
template <bool HBD, int NMAX, int ROWS, class PIX>
__device__ void recon_sb_row(const B200RdoJob &job, const TileCtx &t, ReconWarp<NMAX, PIX> &w, ReconWarp<NMAX, PIX> *luma_w,
                             ReconSync<ROWS> &sy, int role, int slot, int R, int sb_cols, unsigned &sb_seq, unsigned &blk_seq) {
  const B200Frame *f = t.f;
  const int lane = lane_id(), bd = f->bit_depth, stride = t.stride;
  const int np = role == 0 ? 1 : 2, p0 = role == 0 ? 0 : 1;          // planes this warp codes
  const int r0 = t.tr0 + R * 16;                                       // mi row of the superblock row
  const int ring = R % (2 * ROWS), ring_prev = (R + 2 * ROWS - 1) % (2 * ROWS);
  if (lane == 0) atomicExch(&sy.sb_done[role][ring], 0u);
  __syncwarp();
  // compact coefficient stream of this (tile superblock row, plane): starts at the row's first slot of the plane
  const int sbr_frame = r0 >> 4;
  long long cur[2];
  int16_t *cbase[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    cbase[k] = t.coef[p0 + (k < np ? k : 0)] + ((long long)sbr_frame * f->sb_cols + (t.tc0 >> 4)) * B200_SB_COEFS;
    cur[k] = 0;
  }
  const int max_x = f->mi_cols * 4 - 1, max_y = f->mi_rows * 4 - 1;
#pragma unroll 1
  for (int C = 0; C < sb_cols; C++) {
    const int c0 = t.tc0 + C * 16;
    const int sbx = c0 * 4, sby = r0 * 4;
    // luma: its tile buffer of two superblocks ago must have been consumed by the chroma warp (chroma from luma)
    const int tb = role == 0 ? (int)(sb_seq & 1) : 0;
    if (role == 0 && f->planes > 1 && sb_seq >= 2) wait_ge(&sy.uv_sbs[slot], sb_seq - 1);
    // wavefront: the row above must have finished (and written back) the above-right superblock
    if (R > 0) wait_ge(&sy.sb_done[role][ring_prev], (unsigned)min(C + 2, sb_cols));
    // ---- borders: row above (from the frame, written by the row above), left column (kept from the previous superblock)
    {
      const bool have_top = r0 > t.tr0;
#pragma unroll 1
      for (int k = 0; k < np; k++) {
        const PIX *rp = (const PIX *)t.rec[p0 + k];
        if (have_top) {
#pragma unroll 1
          for (int i = lane; i < 130; i += 32) {
            int xx = sbx - 1 + i;
            xx = xx < 0 ? 0 : (xx > max_x ? max_x : xx);
            w.AB[k][i] = __ldcg(rp + (size_t)(sby - 1) * stride + xx);
          }
        }
      }
      __syncwarp();
    }
    PIX *T0 = w.T[role == 0 ? tb : 0], *T1 = w.T[1];
    const int max_lx = max_x - sbx, max_ly = max_y - sby;
    // ---- decided blocks of the superblock in z-order; the next block's mode info and source are fetched ahead
    int z = 0;
    BlockInfo cur_b, nxt_b;
    constexpr int kRounds = NMAX == 32 ? (sizeof(PIX) == 2 ? 4 : 2) : 1;      // 32-lane fetch rounds of the largest block
    SrcRegs<PIX> sreg[2][kRounds];      // [plane k][fetch round]
    auto locate = [&](int zz, int &mr, int &mc) -> bool {      // first coded unit at or after zz
      for (; zz < 256; zz++) {
        int ur = 0, uc = 0;
        for (int b = 0; b < 4; b++) { uc |= ((zz >> (2 * b)) & 1) << b; ur |= ((zz >> (2 * b + 1)) & 1) << b; }
        mr = r0 + ur; mc = c0 + uc;
        if (mr < f->mi_rows && mc < f->mi_cols) { z = zz; return true; }
      }
      z = 256;
      return false;
    };
    auto fetch_info = [&](int mr, int mc, BlockInfo &b) {
      b.m = t.mi[mr * f->mi_cols + mc];
      const B200Mi ma = t.mi[(mr > t.tr0 ? mr - 1 : mr) * f->mi_cols + mc], ml = t.mi[mr * f->mi_cols + (mc > t.tc0 ? mc - 1 : mc)];
      b.mode_above = role == 0 ? ma.y_mode : ma.uv_mode;
      b.mode_left = role == 0 ? ml.y_mode : ml.uv_mode;
    };
    auto fetch_src = [&](int mr, int mc, int n) {
      const int rounds = src_rounds<PIX>(n);
#pragma unroll
      for (int k = 0; k < 2; k++)
#pragma unroll
        for (int h = 0; h < kRounds; h++)
          if (k < np && h < rounds) src_fetch<PIX>((const PIX *)t.src[p0 + k], stride, mc * 4, mr * 4, n, sreg[k][h], h);
    };
    int mr, mc;
    bool have = locate(0, mr, mc);
    if (have) {
      fetch_info(mr, mc, cur_b);
      const int bs = cur_b.m.bsize;
      fetch_src(mr, mc, bs == BLOCK_4X4 ? 4 : (bs == BLOCK_8X8 ? 8 : (bs == BLOCK_16X16 ? 16 : 32)));
    }
#pragma unroll 1
    while (have) {
      const B200Mi m = cur_b.m;
      const int bs = m.bsize;
      const int lg = bs == BLOCK_4X4 ? 2 : (bs == BLOCK_8X8 ? 3 : (bs == BLOCK_16X16 ? 4 : 5));
      const int n = 1 << lg, N2 = n * n, G = 32 >> lg;
      const int x = mc * 4, y = mr * 4, bx = x - sbx, by = y - sby;
      // source block registers -> shared memory (row-major n x n)
      {
        const int rounds = src_rounds<PIX>(n);
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
          for (int h = 0; h < kRounds; h++)
            if (k < np && h < rounds) src_store<PIX>(w.srcb[k], n, sreg[k][h], h);
        __syncwarp();
      }
      // fetch ahead: the next block of this superblock
      int nmr = 0, nmc = 0;
      const int zn = z + (1 << (2 * (lg - 2)));
      const bool have_next = locate(zn, nmr, nmc);
      if (have_next) fetch_info(nmr, nmc, nxt_b);
      const int avail = block_avail(t, x, y, n), ha = avail & 1, hl = (avail >> 1) & 1;
      const bool cfl = role == 1 && m.uv_mode == UV_CFL_PRED;
      int au = 0, av = 0;
      if (cfl) {
        const int js = (m.flags >> 1) & 7, su = (js + 1) / 3, sv = (js + 1) % 3;
        au = su == 0 ? 0 : (su == 1 ? -1 : 1) * ((m.cfl & 15) + 1);
        av = sv == 0 ? 0 : (sv == 1 ? -1 : 1) * ((m.cfl >> 4) + 1);
        // the reconstructed luma of THIS block: wait for the luma warp, then read its tile
        wait_ge(&sy.y_blocks[slot], blk_seq + 1);
        const PIX *LT = luma_w->T[sb_seq & 1] + by * kTS + bx;
        // AC of the reconstructed luma (oracle cfl_ac): (luma << 3) - rounded average
        long long sum = 0;
#pragma unroll 1
        for (int q = lane; q < N2; q += 32) sum += (int)LT[(q >> lg) * kTS + (q & (n - 1))] << 3;
        const int avg = (int)((warp_sum_ll(sum) + (N2 >> 1)) >> (2 * lg));
#pragma unroll 1
        for (int q = lane; q < N2; q += 32) w.ac[q] = (int16_t)(((int)LT[(q >> lg) * kTS + (q & (n - 1))] << 3) - avg);
        __syncwarp();
      }
      // ---- predictions
#pragma unroll 1
      for (int k = 0; k < np; k++) {
        PIX *Tk = k == 0 ? T0 : T1;
        edges_from_tile<PIX>(Tk, w.AB[k], w.LC[k], w.above[k], w.left[k], bx, by, n, avail, bd, max_lx, max_ly);
        PIX *dst = w.pred[k];
        if (cfl) {
          cfl_predict_block(w.ac, N2, dc_value_warp(n, w.above[k] + EOFF, w.left[k] + EOFF, ha, hl, bd), k == 0 ? au : av, bd, dst);
        } else {
          int ief = -1;
          if (f->enable_intra_edge_filter) {      // get_filter_type (spec 7.11.2.8): a smooth mode above or to the left
            ief = 0;
            if (ha && cur_b.mode_above >= SMOOTH_PRED && cur_b.mode_above <= SMOOTH_H_PRED) ief = 1;
            if (hl && cur_b.mode_left >= SMOOTH_PRED && cur_b.mode_left <= SMOOTH_H_PRED) ief = 1;
          }
          const int mode = role == 0 ? m.y_mode : m.uv_mode, delta = (role == 0 ? (m.angle & 15) : (m.angle >> 4)) - 3;
          if (ief >= 0 && mode >= V_PRED && mode <= D67_PRED) predict_block_ief(mode, delta, lg, w.above[k], w.left[k], ha, hl, bd, ief, w.ef[0], w.ef[1], dst);
          else predict_block(mode, delta, lg, w.above[k], w.left[k], ha, hl, bd, dst);
        }
        __syncwarp();
      }
      // fetch ahead: the next block's source rows (the loads are in flight while this block is transformed)
      if (have_next) {
        const int nbs = nxt_b.m.bsize;
        fetch_src(nmr, nmc, nbs == BLOCK_4X4 ? 4 : (nbs == BLOCK_8X8 ? 8 : (nbs == BLOCK_16X16 ? 16 : 32)));
      }
      // ---- transform / quantise / reconstruct: lane group g codes plane p0 + g (32x32: one plane at a time)
      const int tx_type = role == 0 ? m.tx_type : chroma_tx_type(m.uv_mode);
      int eobs[2] = {0, 0};
      if (G >= 2 || np == 1) {
        const int g = lane >> lg;
        const bool active = g < np;
        const int k = active ? g : 0;
        const int e = recon_tx_any<HBD, NMAX, PIX>(n, f, active, w.srcb[k], w.pred[k], (k == 0 ? T0 : T1) + by * kTS + bx,
                                             w.buf + k * n * (n + 1), w.qc[k], tx_type);
        eobs[0] = __shfl_sync(FULL, e, 0);
        if (np == 2) eobs[1] = __shfl_sync(FULL, e, n & 31);
      } else {
#pragma unroll 1
        for (int k = 0; k < 2; k++)
          eobs[k] = recon_tx_any<HBD, NMAX, PIX>(n, f, true, w.srcb[k], w.pred[k], (k == 0 ? T0 : T1) + by * kTS + bx, w.buf, w.qc[k], tx_type);
      }
      // ---- records: eob per plane, levels appended to the compact stream
#pragma unroll 1
      for (int k = 0; k < np; k++) {
        int16_t *dst = cbase[k] + cur[k];
#pragma unroll 1
        for (int i = lane; i < eobs[k]; i += 32) dst[i] = w.qc[k][i];
        cur[k] += eobs[k];
        if (lane == 0) t.eob[p0 + k][mr * f->mi_cols + mc] = (uint16_t)eobs[k];
      }
      if (role == 0) publish_add(&sy.y_blocks[slot]);      // the luma of this block is in the tile (chroma from luma)
      else __syncwarp();
      blk_seq++;
      have = have_next;
      mr = nmr; mc = nmc;
      cur_b = nxt_b;
    }
    // ---- superblock finished: write the tile back (16-byte rows), keep its right column as the next left column
#pragma unroll 1
    for (int k = 0; k < np; k++) {
      const PIX *Tk = k == 0 ? T0 : T1;
      PIX *rp = (PIX *)t.rec[p0 + k];
      const int rows = min(64, max_ly + 1), cols = min(64, max_lx + 1);
      constexpr int per16 = 16 / (int)sizeof(PIX);
      const bool vec = ((stride * (int)sizeof(PIX)) & 15) == 0 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0);
      if (vec) {
        const int segs = (cols + per16 - 1) / per16;      // stride is 16-byte aligned: a row tail never crosses into the next row's samples
#pragma unroll 1
        for (int i = lane; i < rows * segs; i += 32) {
          const int r = i / segs, s = i % segs;
          *reinterpret_cast<uint4 *>(rp + (size_t)(sby + r) * stride + sbx + s * per16) = *reinterpret_cast<const uint4 *>(Tk + r * kTS + s * per16);
        }
      } else {
#pragma unroll 1
        for (int i = lane; i < rows * cols; i += 32) rp[(size_t)(sby + i / cols) * stride + sbx + i % cols] = Tk[(i / cols) * kTS + i % cols];
      }
#pragma unroll 1
      for (int i = lane; i < 64; i += 32) w.LC[k][i] = Tk[i * kTS + 63];
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
      atomicAdd(&sy.sb_done[role][ring], 1u);
      if (role == 1) atomicAdd(&sy.uv_sbs[slot], 1u);
    }
    sb_seq++;
  }
  // stream lengths of this (tile superblock row, plane)
  if (lane == 0)
    for (int k = 0; k < np; k++) job.stream_len[p0 + k][sbr_frame * f->tile_cols + t.tile_col] = (uint32_t)cur[k];
}

}  // namespace

#ifdef B200_EMU
#define B200_RECON_SMEM(TYPE, NAME) static TYPE NAME
#else
#define B200_RECON_SMEM(TYPE, NAME) extern __shared__ __align__(16) unsigned char NAME##_raw[]; TYPE &NAME = *reinterpret_cast<TYPE *>(NAME##_raw)
#endif

template <bool HBD, int NMAX, int ROWS>
struct ReconShared {
  ReconWarp<NMAX, typename PixOf<HBD>::type> w[2 * ROWS];     // [slot * 2 + role]
  ReconSync<ROWS> sy;
};

// Register budget: the chain is latency-bound, so what matters is that EVERY tile of a launch is resident at once (one
// wave) while leaving room for the throughput kernels of the next chunk on the same SMs: <= 128 registers per thread
// (16x16 transforms hold 3 x 16 values per lane); the 32x32 instances need the full budget.
template <bool HBD, int NMAX, int ROWS>
__global__ void __launch_bounds__(64 * ROWS, NMAX == 32 ? 1 : (ROWS == 1 ? 8 : (ROWS == 2 ? 4 : 2)))  /* 128 registers per thread */ b200_recon_kernel(const B200RdoJob *jobs, const B200TileRef *tiles) {
  typedef typename PixOf<HBD>::type PIX;
  typedef ReconShared<HBD, NMAX, ROWS> SH;
  B200_RECON_SMEM(SH, sh);
  const B200TileRef tr = tiles[blockIdx.x];
  const B200RdoJob &job = jobs[tr.job];
  TileCtx t;
  ctx_from_job(job, t);
  const B200Frame *f = t.f;
  t.tr0 = f->tile_row_start_sb[tr.tile_row] * 16;
  t.tr1 = min(f->tile_row_start_sb[tr.tile_row + 1] * 16, f->mi_rows);
  t.tc0 = f->tile_col_start_sb[tr.tile_col] * 16;
  t.tc1 = min(f->tile_col_start_sb[tr.tile_col + 1] * 16, f->mi_cols);
  t.tile_col = tr.tile_col;
  const int sb_rows = (t.tr1 - t.tr0 + 15) >> 4, sb_cols = (t.tc1 - t.tc0 + 15) >> 4;
  for (int i = threadIdx.x; i < (int)(sizeof(ReconSync<ROWS>) / 4); i += blockDim.x) reinterpret_cast<unsigned *>(&sh.sy)[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, slot = warp >> 1, role = warp & 1;
  if (!(role == 1 && f->planes == 1)) {
    unsigned sb_seq = 0, blk_seq = 0;
#pragma unroll 1
    for (int R = slot; R < sb_rows; R += ROWS)
      recon_sb_row<HBD, NMAX, ROWS, PIX>(job, t, sh.w[slot * 2 + role], &sh.w[slot * 2], sh.sy, role, slot, R, sb_cols, sb_seq, blk_seq);
  }
  __syncthreads();
  // ---- skip flags: a block whose planes all quantised to zero is coded as "skip" (every unit of the block carries the flag)
  const int tw = t.tc1 - t.tc0, units = (t.tr1 - t.tr0) * tw;
  for (int i = threadIdx.x; i < units; i += blockDim.x) {
    const int mr = t.tr0 + i / tw, mc = t.tc0 + i % tw;
    B200Mi *mp = &t.mi[mr * f->mi_cols + mc];
    const int bs = mp->bsize;
    const int n4 = bs == BLOCK_4X4 ? 1 : (bs == BLOCK_8X8 ? 2 : (bs == BLOCK_16X16 ? 4 : 8));
    const int orow = mr & ~(n4 - 1), ocol = mc & ~(n4 - 1);
    bool any = false;
    for (int p = 0; p < f->planes; p++) any |= t.eob[p][orow * f->mi_cols + ocol] != 0;
    if (!any) mp->flags |= 1;
  }
}

// ---- gather: the compact streams of all (job, plane, tile superblock row) regions -> one contiguous array ------------
// regions[i]: {source = device pointer of the region's stream, length pointer}; offsets by a single-CTA exclusive scan.
struct B200StreamRegion { const int16_t *src; const uint32_t *len; };

__global__ void __launch_bounds__(1024) b200_stream_scan_kernel(const B200StreamRegion *regions, int n, uint32_t *offsets /* n + 1 */) {
  __shared__ uint32_t part[1024];
  const int tid = threadIdx.x, per = (n + 1023) / 1024;
  uint32_t s = 0;
  for (int i = tid * per; i < min((tid + 1) * per, n); i++) s += *regions[i].len;
  part[tid] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const uint32_t v = tid >= d ? part[tid - d] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  uint32_t run = tid ? part[tid - 1] : 0;
  for (int i = tid * per; i < min((tid + 1) * per, n); i++) { offsets[i] = run; run += *regions[i].len; }
  if (tid == 1023) offsets[n] = part[1023];
}
// one CTA per region: copy its levels to out + offsets[region] (out may be mapped pinned host memory)
__global__ void __launch_bounds__(256) b200_stream_gather_kernel(const B200StreamRegion *regions, const uint32_t *offsets, int16_t *out, uint32_t capacity) {
  const B200StreamRegion rg = regions[blockIdx.x];
  const uint32_t len = *rg.len, off = offsets[blockIdx.x];
  if (off + len > capacity) return;            // the host sees offsets[n] > capacity and reports the overflow
  for (uint32_t i = threadIdx.x; i < len; i += 256) out[off + i] = rg.src[i];
}

#ifndef B200_EMU
namespace {
template <bool HBD, int NMAX, int ROWS>
int recon_launch_t(const B200RdoJob *d_jobs, const B200TileRef *d_tiles, int n_tiles, cudaStream_t st) {
  typedef ReconShared<HBD, NMAX, ROWS> SH;
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!done[dev & 63]) {
      cudaFuncSetAttribute(b200_recon_kernel<HBD, NMAX, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SH));
      cudaFuncSetAttribute(b200_recon_kernel<HBD, NMAX, ROWS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      done[dev & 63] = true;
    }
  }
  b200_recon_kernel<HBD, NMAX, ROWS><<<n_tiles, 64 * ROWS, sizeof(SH), st>>>(d_jobs, d_tiles);
  return (int)cudaGetLastError();
}
}  // namespace

// rows: superblock-row slots per tile (1 .. 4: more slots = deeper wavefront, more shared memory per tile)
extern "C" int b200_recon_launch(const B200RdoJob *d_jobs, const B200TileRef *d_tiles, int n_tiles, int high_bitdepth, int top_level, int rows,
                                 void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_tiles <= 0) return 0;
#define B200_RECON_GO(HBD, NMAX) \
  (rows >= 4 ? recon_launch_t<HBD, NMAX, 4>(d_jobs, d_tiles, n_tiles, st) : rows == 3 ? recon_launch_t<HBD, NMAX, 3>(d_jobs, d_tiles, n_tiles, st) \
   : rows == 2 ? recon_launch_t<HBD, NMAX, 2>(d_jobs, d_tiles, n_tiles, st) : recon_launch_t<HBD, NMAX, 1>(d_jobs, d_tiles, n_tiles, st))
  if (top_level >= 3) return high_bitdepth ? B200_RECON_GO(true, 32) : B200_RECON_GO(false, 32);
  return high_bitdepth ? B200_RECON_GO(true, 16) : B200_RECON_GO(false, 16);
}
extern "C" size_t b200_recon_smem_bytes(int high_bitdepth, int top_level, int rows) {
#define B200_RECON_SZ(HBD, NMAX) (rows >= 4 ? sizeof(ReconShared<HBD, NMAX, 4>) : rows == 3 ? sizeof(ReconShared<HBD, NMAX, 3>) : rows == 2 ? sizeof(ReconShared<HBD, NMAX, 2>) : sizeof(ReconShared<HBD, NMAX, 1>))
  if (top_level >= 3) return high_bitdepth ? B200_RECON_SZ(true, 32) : B200_RECON_SZ(false, 32);
  return high_bitdepth ? B200_RECON_SZ(true, 16) : B200_RECON_SZ(false, 16);
}
extern "C" int b200_stream_gather_launch(const void *d_regions, int n_regions, uint32_t *d_offsets, int16_t *out, uint32_t capacity, void *stream) {
  if (n_regions <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  b200_stream_scan_kernel<<<1, 1024, 0, st>>>((const B200StreamRegion *)d_regions, n_regions, d_offsets);
  b200_stream_gather_kernel<<<n_regions, 256, 0, st>>>((const B200StreamRegion *)d_regions, d_offsets, out, capacity);
  B200_MARK("stream_gather", st);
  return (int)cudaGetLastError();
}
#else
static const B200RdoJob *g_rc_jobs;
static const B200TileRef *g_rc_tiles;
static int g_rc_top;
template <int ROWS> static void emu_recon_entry(void *) {
  const bool hbd = g_rc_jobs[0].frame.bit_depth > 8;
  if (g_rc_top == 3) { if (hbd) b200_recon_kernel<true, 32, ROWS>(g_rc_jobs, g_rc_tiles); else b200_recon_kernel<false, 32, ROWS>(g_rc_jobs, g_rc_tiles); }
  else { if (hbd) b200_recon_kernel<true, 16, ROWS>(g_rc_jobs, g_rc_tiles); else b200_recon_kernel<false, 16, ROWS>(g_rc_jobs, g_rc_tiles); }
}
extern "C" int b200_recon_launch_emu(const B200RdoJob *jobs, const B200TileRef *tiles, int n_tiles, int rows) {
  g_rc_jobs = jobs; g_rc_tiles = tiles;
  g_rc_top = jobs[0].frame.partition_max >= 32 ? 3 : 2;
  if (rows >= 4) emu::launch(emu_recon_entry<4>, nullptr, (unsigned)n_tiles, 64 * 4);
  else if (rows == 3) emu::launch(emu_recon_entry<3>, nullptr, (unsigned)n_tiles, 64 * 3);
  else if (rows >= 2) emu::launch(emu_recon_entry<2>, nullptr, (unsigned)n_tiles, 64 * 2);
  else emu::launch(emu_recon_entry<1>, nullptr, (unsigned)n_tiles, 64);
  return 0;
}
#endif
