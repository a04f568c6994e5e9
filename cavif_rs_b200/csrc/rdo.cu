// AV1 intra rate-distortion search on the device, in three stages:
//
//   A. ANALYSIS (open loop, embarrassingly parallel).  Every block of every level (4x4, 8x8, 16x16, 32x32 when the
//      partition range reaches it) of every image is analysed independently, predicting from SOURCE neighbours, from
//      the largest level down so that a level can be pruned by its parent (b200_pruned_by_parent, records.h):
//      luma SATD ranking of the 13 intra modes + angle refinement (b200_satd_direct_kernel for the never-pruned top
//      level, b200_satd_kernel with per-CTA compaction of the surviving blocks below it: one thread per block), then
//      full trials of the top-2 (top-4: frame.luma_rdo_modes) luma modes x transform types and of the chroma candidates
//      {luma mode, DC, CFL}
//      (b200_analyze_kernel: one warp per block, lane-batched candidates: lane group g = lane / N owns one candidate,
//      lane l = lane % N owns column l / row l of its transform; butterflies in registers, transpose through shared
//      memory, quantiser / rate / dequantiser on the row, SSE / Psychovisual sums by shuffles; the warps of a CTA run in
//      lock-step for instruction-cache locality: the top level straight over its blocks, the pruned levels over the
//      survivor lists the SATD kernel built).
//   B. PARTITION DECISION (b200_decide_kernel: one thread per root region, NONE vs SPLIT recursively).
//   C. RECONSTRUCTION (closed loop, recon.cu b200_recon_kernel: one CTA per TILE, warps on a superblock-row wavefront): the only
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
#include "av1_dev.cuh"
#ifndef B200_EMU
#include <mutex>
#endif

namespace {

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
  typedef typename SH::Pix PIX;
  if (SYNC) __syncthreads();
  const B200Frame *f = t.f;
  const B200RateTables *rt = t.rt;
  const int lane = lane_id(), g = lane >> lg;
  const int planes = f->planes, bd = f->bit_depth;
  // ---- stage source block and SOURCE edges; clear the padded level maps
  {   // one 4 / 8 / 16-byte row segment per lane: global -> registers -> shared memory (n x n row-major)
    constexpr int kRowBytes = n * (int)sizeof(PIX), kSeg = kRowBytes < 16 ? kRowBytes : 16, kRounds = (n * (kRowBytes / kSeg) + 31) / 32;
    SrcRegs<PIX> sr[3][kRounds];
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
      for (int h = 0; h < kRounds; h++)
        if (p < planes) src_fetch<PIX>((const PIX *)t.src[p], t.stride, x, y, n, sr[p][h], h);
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
      for (int h = 0; h < kRounds; h++)
        if (p < planes) src_store<PIX>(sh.src[p], n, sr[p][h], h);
  }
#pragma unroll 1
  for (int i = lane; i < (int)sizeof(sh.lev) / 16; i += 32) reinterpret_cast<uint4 *>(sh.lev)[i] = make_uint4(0, 0, 0, 0);
  const int avail = block_avail(t, x, y, n), ha = avail & 1, hl = (avail >> 1) & 1;
#pragma unroll 1
  for (int p = 0; p < planes; p++) build_edges(t, sh, (const PIX *)t.src[p], p, x, y, n, avail);
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
  int r0, r1, r2, r3;          // ranked modes; full trials on the first frame.luma_rdo_modes (2 or 4) of them, pair by pair
  rank_top4(my_ycost, r0, r1, r2, r3);
  const int n_pairs = f->luma_rdo_modes >= 4 ? 2 : 1;
  // ---- luma full trials: candidate c = k * B200_TX_TYPES + ti; lane batches pair the two modes on the same transform type
  const int tx_mask = b200_tx_mask_for(f->tx_type_mask, f->reduced_tx_set, n);      // 32x32: DCT only; 1-D DCTs: 7-type set only
  int first_t = -1;
  for (int k = 0; k < B200_TX_TYPES; k++) if ((tx_mask >> k) & 1) { first_t = k; break; }
  // open loop: the neighbours' modes are not decided yet, the trials assume filter type 0 (oracle)
  const int ief = f->enable_intra_edge_filter ? 0 : -1;
  long long bestJ = kInf;
  int bestC = 1 << 20, bestEob = 0;
#pragma unroll 1
  for (int pr = 0; pr < n_pairs; pr++) {
  const int m0 = pr == 0 ? r0 : r2, m1 = pr == 0 ? r1 : r3;
  __syncwarp();
  predict_any(sh, ief, m0 >= 0 ? m0 : 0, DELTA_Y(m0 >= 0 ? m0 : 0), lg, sh.above[0], sh.left[0], ha, hl, bd, sh.pred);
  predict_any(sh, ief, m1 >= 0 ? m1 : (m0 >= 0 ? m0 : 0), DELTA_Y(m1 >= 0 ? m1 : (m0 >= 0 ? m0 : 0)), lg, sh.above[0], sh.left[0], ha, hl, bd, sh.pred + PS);
  __syncwarp();
  bool skip_rest = false;      // every first-type trial of the pair quantised to all-zero: its other transform types are not tried
#pragma unroll 1
  for (int b = 0; b * G < 2 * B200_TX_TYPES && (SYNC || !skip_rest); b++) {
    if (SYNC) { __syncthreads(); if (skip_rest) continue; }
    const int ci = b * G + g;
    const int ti = ci >> 1, k = ci & 1;
    const int m = k == 0 ? m0 : m1;
    bool ok = ti < B200_TX_TYPES && m >= 0 && ((tx_mask >> ti) & 1) && (f->rdo_tx_decision || ti == first_t);
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
    if (ok) { wj = rd_cost(f, r.sse, r.rate + mode_rate(rt, ymode_cost, m, n, DELTA_Y(m))); wc = (2 * pr + k) * B200_TX_TYPES + ti; }
    warp_argmin(wj, wc);
    if (wj < bestJ || (wj == bestJ && wc < bestC)) {
      bestJ = wj; bestC = wc;
      const int wg = (wc % B200_TX_TYPES) * 2 + (wc / B200_TX_TYPES - 2 * pr) - b * G;
      bestEob = __shfl_sync(FULL, r.eob, wg << lg);
    }
  }
  }
  const int best_rank = bestC / B200_TX_TYPES;
  const int best_mode = best_rank == 0 ? r0 : (best_rank == 1 ? r1 : (best_rank == 2 ? r2 : r3));
  const int best_tt = dv_tx_types[bestC % B200_TX_TYPES];
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
        PIX *dst = sh.pred + (cc - b * G) * N2;
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

}  // namespace

// ---- stage A.1: SATD of the 13 intra modes for every block of level lv, predicted from SOURCE neighbours.
// One thread per (block, plane group): group 0 = luma, group 1 = U + V.  out: [block][2][13] uint32.
// One item = one block of level lv (luma SATD of the 13 modes + angle refinement).  A CTA owns kSatdItems consecutive items: it first compacts the items
// that survive the top-down pruning into shared memory, then its 128 threads walk the compacted list, so warps stay
// full even when most of a level is pruned.
constexpr int kSatdItems = 512;
template <class PIX> __device__ __forceinline__ void satd_item(const B200RdoJob &job, int lv, int gid, int *sh_d);

template <class PIX>
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
      if (!active) {                       // never split down to this block
        B200BlockDec d;
        d.cost = B200_COST_PRUNED; d.y_mode = 0; d.tx_type = 0; d.uv_mode = 0; d.all_zero = 0; d.angle_y = 0; d.angle_uv = 0;
        d.cfl_js = 0; d.cfl_alpha = 0;
        job.dec[lv][blk] = d;
      }
    }
    const unsigned m = __ballot_sync(FULL, active);
    int off = 0;
    if (lane_id() == 0 && m) off = atomicAdd(&s_n, __popc(m));
    off = __shfl_sync(FULL, off, 0);
    if (active) s_list[off + __popc(m & ((1u << lane_id()) - 1))] = (uint16_t)it;
  }
  __syncthreads();
  const int cnt = s_n;
  // the survivors also go to the level's global list (walked by the lock-step analysis kernel)
  __shared__ unsigned s_goff;
  if (threadIdx.x == 0 && cnt) s_goff = atomicAdd(&job.list_n[lv], (unsigned)cnt);
  __syncthreads();
#pragma unroll 1
  for (int i = threadIdx.x; i < cnt; i += 128) job.list[lv][s_goff + i] = (uint32_t)(base + s_list[i]);
#pragma unroll 1
  for (int i = threadIdx.x; i < cnt; i += 128) satd_item<PIX>(job, lv, base + s_list[i], sh_d);
}

// first launch of the stage: empty survivor lists for every job of the launch
__global__ void b200_list_reset_kernel(const B200RdoJob *jobs, int n_jobs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_jobs * B200_LEVELS && jobs[i / B200_LEVELS].list_n) jobs[i / B200_LEVELS].list_n[i % B200_LEVELS] = 0;
}

// never-pruned top level: one item per thread, no compaction
template <class PIX>
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
  satd_item<PIX>(job, lv, gid, sh_d);
}

// Prediction of the 4x4 sub-block at (sy, sx) of an n x n block, mode m / angle delta, written to out[0..15] (row-major).
// Same values as pred_px() pixel by pixel; the mode is resolved once per sub-block and every class walks the 16 pixels
// with its own short loop (pred_px() re-derives the angle, the zone and the table entries for every pixel).
__device__ __noinline__ void satd_pred4x4(int m, int delta, int n, int sy, int sx, const uint16_t *A, const uint16_t *L, int dc, int *out) {
  if (m == DC_PRED) {
#pragma unroll 1
    for (int k = 0; k < 16; k++) out[k] = dc;
    return;
  }
  if (m == PAETH_PRED) {
    const int tl = A[-1];
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      const int l = L[sy + i];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int a = A[sx + j], b = a + l - tl;
        const int pl = abs(b - l), pt = abs(b - a), ptl = abs(b - tl);
        out[i * 4 + j] = (pl <= pt && pl <= ptl) ? l : (pt <= ptl ? a : tl);
      }
    }
    return;
  }
  if (m >= SMOOTH_PRED) {
    const uint8_t *w = dv_sm_weights + (n == 4 ? 0 : n == 8 ? 4 : n == 16 ? 12 : 28);
    const int bl = L[n - 1], tr = A[n - 1];
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      const int wi = w[sy + i], l = L[sy + i];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int wj = w[sx + j];
        const int v = wi * A[sx + j] + (256 - wi) * bl, h = wj * l + (256 - wj) * tr;
        out[i * 4 + j] = m == SMOOTH_PRED ? (v + h + 256) >> 9 : (m == SMOOTH_V_PRED ? (v + 128) >> 8 : (h + 128) >> 8);
      }
    }
    return;
  }
  const int p_angle = dv_mode_to_angle[m] + 3 * delta;
  if (p_angle == 90 || p_angle == 180) {
#pragma unroll 1
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) out[i * 4 + j] = p_angle == 90 ? A[sx + j] : L[sy + i];
    return;
  }
  if (p_angle > 90 && p_angle < 180) {
    const int dx = dv_dr_intra_derivative[180 - p_angle], dy = dv_dr_intra_derivative[p_angle - 90];
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      const int ii = sy + i, rowx = (ii + 1) * dx;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int jj = sx + j;
        int idx = (jj << 6) - rowx, base = idx >> 6;
        const uint16_t *e = A;
        if (base < -1) { idx = (ii << 6) - (jj + 1) * dy; base = idx >> 6; e = L; }
        const int shift = (idx >> 1) & 31;
        out[i * 4 + j] = (e[base] * (32 - shift) + e[base + 1] * shift + 16) >> 5;
      }
    }
    return;
  }
  // zone 1 (above row, the walk depends on the pixel row) / zone 3 (left column, it depends on the pixel column)
  const bool z1 = p_angle < 90;
  const uint16_t *e = z1 ? A : L;
  const int d = dv_dr_intra_derivative[z1 ? p_angle : 270 - p_angle], max_base = 2 * n - 1;
  const int u0 = z1 ? sy : sx, v0 = z1 ? sx : sy;     // u: the coordinate the walk depends on, v: the other one
#pragma unroll 1
  for (int a = 0; a < 4; a++) {
    const int idx = (u0 + a + 1) * d, b0 = (idx >> 6) + v0, shift = (idx >> 1) & 31;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int base = b0 + b;
      const int v = base < max_base ? (e[base] * (32 - shift) + e[base + 1] * shift + 16) >> 5 : e[max_base];
      out[z1 ? a * 4 + b : b * 4 + a] = v;
    }
  }
}

template <class PIX>
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
      const PIX *r = (const PIX *)job.src[p];
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
          satd_pred4x4(m, delta, n, sy, sx, A, L, dc, sh_d + threadIdx.x * 17);
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
#ifndef B200_LV2_WARPS
#define B200_LV2_WARPS 20     // resident warps per SM the 16x16 analysis kernels are compiled for (register budget 65536 / 32 / warps)
#endif
#ifndef B200_LV01_WARPS
#define B200_LV01_WARPS 24    // resident warps per SM of the 8x8 / 4x4 analysis kernels (80 registers)
#endif
#ifndef B200_TOP_W
#define B200_TOP_W 5
#endif
constexpr int kAnalyzeWarpsTop = B200_TOP_W;      // 16x16 as the top level
#ifndef B200_LISTED_W
#define B200_LISTED_W 4
#endif
constexpr int kListedW = B200_LISTED_W;      // warps per lock-step CTA of the pruned levels
constexpr int kAnalyzeWarpsTop32 = 2;    // 32x32 as the top level (22.9 KB of shared memory per warp)
template <int LV, bool HBD, int W>
__global__ void __launch_bounds__(kThreads * W, W == 6 ? 3 : (LV == 3 ? 4 : (LV == 2 ? B200_LV2_WARPS : 24) / W)) b200_analyze_kernel(const B200RdoJob *jobs) {
  typedef typename PixOf<HBD>::type PIX;
  typedef SharedT<(LV == 3 ? 32 : 16), PIX> SH;
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
          for (int i = lane_id(); i < n * f->planes; i += 32) B200_PREFETCH_L2((const PIX *)t.src[i / n] + (size_t)(ny + i % n) * t.stride + nx);
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

// Pruned levels: the W warps of a CTA take W consecutive entries of the level's SURVIVOR LIST (built by the SATD kernel) and
// analyse them in lock-step, exactly like the top level - no pruned block is ever visited, no warp idles inside a segment, and
// the warps of a CTA share one instruction stream (free-running warps lose most issue slots to instruction fetch here:
// stall_no_instruction 6-10 warps per issue, profiles/).  CTAs walk the list grid-stride.
template <int LV, bool HBD, int W>
__global__ void __launch_bounds__(kThreads * W, (LV == 2 ? B200_LV2_WARPS : B200_LV01_WARPS) / W) b200_analyze_listed_kernel(const B200RdoJob *jobs) {
  typedef typename PixOf<HBD>::type PIX;
  typedef SharedT<16, PIX> SH;
  __shared__ SH shs[W];
  SH &sh = shs[threadIdx.x >> 5];
  const B200RdoJob &job = jobs[blockIdx.y];
  const B200Frame *f = &job.frame;
  constexpr int lv = LV, n = 4 << LV;
  if (n < min(f->partition_min, 8) || n > f->partition_max) return;
  TileCtx t;
  ctx_from_job(job, t);
  const int cols = b200_level_cols(f, lv);
  const int total = (int)job.list_n[lv];
  const int per_iter = gridDim.x * W, iters = (total + per_iter - 1) / per_iter;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    const int want = (it * gridDim.x + blockIdx.x) * W + (threadIdx.x >> 5);
    const int blk = (int)job.list[lv][want < total ? want : total - 1];        // padding warps repeat the last entry, output discarded
    const int x = (blk % cols) * n, y = (blk / cols) * n;
    tile_of_unit(f, y >> 2, x >> 2, t);
    analyze_block<2 + LV, HBD, true>(t, sh, x, y, job.satd[lv] + (size_t)blk * B200_SATD_WORDS, want < total ? job.dec[lv] + blk : nullptr);
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

#ifndef B200_EMU
// Stages A + B go to `stream`; stage C goes to `stream_recon` (may be the same stream) after `analysis_done` (an event
// recorded behind stage B, may be null when both streams are the same).  The caller overlaps the low-parallelism
// reconstruction chain of one chunk with the analysis of the next by giving stream_recon a higher priority.
extern "C" int b200_rdo_launch(const B200RdoJob *d_jobs, int n_jobs, const B200TileRef *d_tiles, int n_tiles, int max_blocks4,
                               int max_regions, int high_bitdepth, int top_level, int recon_rows, void *stream, void *stream_recon, void *analysis_done, int *launched) {
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
    B200_CARVE((b200_analyze_kernel<2, false, kAnalyzeWarpsTop>)); B200_CARVE((b200_analyze_kernel<2, true, kAnalyzeWarpsTop>));
    B200_CARVE((b200_analyze_kernel<2, false, 6>));
    B200_CARVE((b200_analyze_listed_kernel<0, false, kListedW>)); B200_CARVE((b200_analyze_listed_kernel<1, false, kListedW>)); B200_CARVE((b200_analyze_listed_kernel<2, false, kListedW>));
    B200_CARVE((b200_analyze_listed_kernel<0, true, kListedW>)); B200_CARVE((b200_analyze_listed_kernel<1, true, kListedW>)); B200_CARVE((b200_analyze_listed_kernel<2, true, kListedW>));
    B200_CARVE((b200_analyze_kernel<3, false, kAnalyzeWarpsTop32>)); B200_CARVE((b200_analyze_kernel<3, true, kAnalyzeWarpsTop32>));
    B200_CARVE((b200_analyze_kernel<0, false, 1>)); B200_CARVE((b200_analyze_kernel<1, false, 1>)); B200_CARVE((b200_analyze_kernel<2, false, 1>));
    B200_CARVE((b200_analyze_kernel<0, true, 1>)); B200_CARVE((b200_analyze_kernel<1, true, 1>)); B200_CARVE((b200_analyze_kernel<2, true, 1>));
    configured = true;
  }
  const int top = top_level >= 3 ? 3 : 2;
  b200_list_reset_kernel<<<(n_jobs * B200_LEVELS + 127) / 128, 128, 0, st>>>(d_jobs, n_jobs);
  n_launch++;
  B200_MARK(nullptr, st);
  for (int lv = top; lv >= 0; lv--) {   // large to small: every level is pruned by its parent level
    const int nblk = max_blocks4 >> (2 * lv);
    if (nblk == 0) continue;
    // the top level is never pruned: one item per thread; lower levels are mostly pruned: compact 512 items per CTA
    if (lv == top) {
      dim3 gs((nblk + 127) / 128, n_jobs);
      if (high_bitdepth) b200_satd_direct_kernel<uint16_t><<<gs, 128, 0, st>>>(d_jobs, lv);
      else b200_satd_direct_kernel<uint8_t><<<gs, 128, 0, st>>>(d_jobs, lv);
      B200_MARK(lv == 3 ? "satd_32x32" : "satd_16x16", st);
      n_launch++;
    } else {
      dim3 gs((nblk + kSatdItems - 1) / kSatdItems, n_jobs);
      if (high_bitdepth) b200_satd_kernel<uint16_t><<<gs, 128, 0, st>>>(d_jobs, lv, kSatdItems);
      else b200_satd_kernel<uint8_t><<<gs, 128, 0, st>>>(d_jobs, lv, kSatdItems);
      B200_MARK(lv == 2 ? "satd_16x16" : (lv == 1 ? "satd_8x8" : "satd_4x4"), st);
      n_launch++;
    }
    // enough warps to fill the machine several times over, grid-stride inside
    const int per_job = min(nblk, max(1, (148 * (lv == 2 ? B200_LV2_WARPS : 16) * 4) / n_jobs));
    dim3 ga(per_job, n_jobs);
#define B200_ANALYZE(LV, W, GRID) \
    do { if (high_bitdepth) b200_analyze_kernel<LV, true, W><<<GRID, kThreads * W, 0, st>>>(d_jobs); \
         else b200_analyze_kernel<LV, false, W><<<GRID, kThreads * W, 0, st>>>(d_jobs); } while (0)
    // never-pruned top level, lock-step CTAs of W warps that each walk ceil(blocks / (CTAs * W)) blocks: B200_TOP_WAVES waves
    // of resident CTAs over the launch, so that the last grid-stride iteration is (nearly) full - with many short waves
    // the partly filled last iteration of every CTA was up to 8 % of the kernel
#ifndef B200_TOP_WAVES
#define B200_TOP_WAVES 2
#endif
    auto top_grid = [&](int warps_per_cta, int ctas_per_sm) {
      const int per_job_ctas = std::max(1, (148 * ctas_per_sm * B200_TOP_WAVES) / n_jobs);
      return dim3((unsigned)std::min(per_job_ctas, (nblk + warps_per_cta - 1) / warps_per_cta), (unsigned)n_jobs);
    };
    if (lv == 3) {
      B200_ANALYZE(3, kAnalyzeWarpsTop32, top_grid(kAnalyzeWarpsTop32, 4));
    } else if (lv == 2 && top == 2) {
      static const int top_warps = getenv("B200AVIF_TOP_WARPS") ? atoi(getenv("B200AVIF_TOP_WARPS")) : kAnalyzeWarpsTop;      // experiment knob
      if (top_warps == 6 && !high_bitdepth) {
        b200_analyze_kernel<2, false, 6><<<top_grid(6, 3), kThreads * 6, 0, st>>>(d_jobs);
      } else {
        B200_ANALYZE(2, kAnalyzeWarpsTop, top_grid(kAnalyzeWarpsTop, B200_LV2_WARPS / kAnalyzeWarpsTop));
      }
    } else {
      static const bool free_running = getenv("B200AVIF_FREE_RUNNING") != nullptr;      // A/B knob: the single-warp kernels of round 1
      if (free_running) {
        if (lv == 2) B200_ANALYZE(2, 1, ga); else if (lv == 1) B200_ANALYZE(1, 1, ga); else B200_ANALYZE(0, 1, ga);
      } else {
        dim3 gl(min((nblk + kListedW - 1) / kListedW, max(1, (148 * ((lv == 2 ? B200_LV2_WARPS : B200_LV01_WARPS) / kListedW) * 2) / n_jobs)), n_jobs);      // two waves of resident CTAs
#define B200_ANALYZE_LISTED(LV) \
        do { if (high_bitdepth) b200_analyze_listed_kernel<LV, true, kListedW><<<gl, kThreads * kListedW, 0, st>>>(d_jobs); \
             else b200_analyze_listed_kernel<LV, false, kListedW><<<gl, kThreads * kListedW, 0, st>>>(d_jobs); } while (0)
        if (lv == 2) B200_ANALYZE_LISTED(2); else if (lv == 1) B200_ANALYZE_LISTED(1); else B200_ANALYZE_LISTED(0);
      }
    }
    B200_MARK(lv == 3 ? "analyze_32x32" : (lv == 2 ? "analyze_16x16" : (lv == 1 ? "analyze_8x8" : "analyze_4x4")), st);
    n_launch++;
  }
  // one thread per ROOT REGION: ceil(W / rn) * ceil(H / rn) per job (larger than n_mi / 16 when the coded size is
  // 8 mod 16), computed by the caller over all jobs
  dim3 gd((max_regions + 127) / 128, n_jobs);
  b200_decide_kernel<<<gd, 128, 0, st>>>(d_jobs);
  B200_MARK("decide", st);
  n_launch += 2;      // partition decision + the reconstruction chain below
  cudaStream_t sr = stream_recon ? (cudaStream_t)stream_recon : st;
  if (analysis_done) {
    cudaEventRecord((cudaEvent_t)analysis_done, st);
    if (sr != st) cudaStreamWaitEvent(sr, (cudaEvent_t)analysis_done, 0);
  }
  {
    B200_MARK(nullptr, sr);
    const int re = b200_recon_launch(d_jobs, d_tiles, n_tiles, high_bitdepth, top, recon_rows, sr);
    B200_MARK("recon_chain", sr);
    if (re) return re;
  }
  if (launched) *launched = n_launch;
  return (int)cudaGetLastError();
}
#else
static const B200RdoJob *g_emu_jobs;
static const B200TileRef *g_emu_tiles;
static int g_emu_lv, g_emu_top;
static void emu_satd(void *) {
  const bool hbd = g_emu_jobs[0].frame.bit_depth > 8;
  if (g_emu_lv == g_emu_top) { if (hbd) b200_satd_direct_kernel<uint16_t>(g_emu_jobs, g_emu_lv); else b200_satd_direct_kernel<uint8_t>(g_emu_jobs, g_emu_lv); }
  else { if (hbd) b200_satd_kernel<uint16_t>(g_emu_jobs, g_emu_lv, kSatdItems); else b200_satd_kernel<uint8_t>(g_emu_jobs, g_emu_lv, kSatdItems); }
}
static void emu_analyze(void *) {
  const bool hbd = g_emu_jobs[0].frame.bit_depth > 8;
#define EMU_ANALYZE(LV, W) do { if (hbd) b200_analyze_kernel<LV, true, W>(g_emu_jobs); else b200_analyze_kernel<LV, false, W>(g_emu_jobs); } while (0)
#define EMU_LISTED(LV) do { if (hbd) b200_analyze_listed_kernel<LV, true, kListedW>(g_emu_jobs); else b200_analyze_listed_kernel<LV, false, kListedW>(g_emu_jobs); } while (0)
  if (g_emu_lv == g_emu_top) { if (g_emu_top == 3) EMU_ANALYZE(3, kAnalyzeWarpsTop32); else EMU_ANALYZE(2, kAnalyzeWarpsTop); }
  else if (g_emu_lv == 2) EMU_LISTED(2);
  else if (g_emu_lv == 1) EMU_LISTED(1);
  else EMU_LISTED(0);
}
static void emu_decide(void *) { b200_decide_kernel(g_emu_jobs); }
extern "C" int b200_recon_launch_emu(const B200RdoJob *jobs, const B200TileRef *tiles, int n_tiles, int rows);
extern "C" int b200_rdo_launch_emu(const B200RdoJob *jobs, const B200TileRef *tiles, int n_tiles, int recon_rows) {
  g_emu_jobs = jobs; g_emu_tiles = tiles;
  const B200Frame *f = &jobs[0].frame;
  g_emu_top = f->partition_max >= 32 ? 3 : 2;
  if (jobs[0].list_n) for (int k = 0; k < B200_LEVELS; k++) jobs[0].list_n[k] = 0;
  for (g_emu_lv = g_emu_top; g_emu_lv >= 0; g_emu_lv--) {
    const int nblk = (f->mi_rows >> g_emu_lv) * (f->mi_cols >> g_emu_lv);
    if (nblk == 0) continue;
    const int items = g_emu_lv == g_emu_top ? 128 : kSatdItems;
    emu::launch(emu_satd, nullptr, (unsigned)((nblk + items - 1) / items), 128);
    const int w = g_emu_lv == g_emu_top ? (g_emu_top == 3 ? kAnalyzeWarpsTop32 : kAnalyzeWarpsTop) : kListedW;
    emu::launch(emu_analyze, nullptr, (unsigned)std::min(nblk, w > 1 ? 3 : 7), kThreads * w);
  }
  const int rn4 = g_emu_top == 3 ? 8 : 4;
  const int regions = ((f->mi_cols + rn4 - 1) / rn4) * ((f->mi_rows + rn4 - 1) / rn4);
  emu::launch(emu_decide, nullptr, (unsigned)((regions + 127) / 128), 128);
  b200_recon_launch_emu(jobs, tiles, n_tiles, recon_rows);
  return 0;
}
#endif
