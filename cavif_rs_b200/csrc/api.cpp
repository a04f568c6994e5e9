// C ABI + host pipeline of the B200 AVIF encode path (include/b200avif.h).
//
// Host-side mirror of ravif::Encoder (/root/reference/ravif/src/av1encoder.rs:66-481): the
// builder state, encode_rgba / encode_rgb / encode_raw_planes_* flow, the colour||alpha split into
// two independent AV1 encodes (:400-481) and the call into the muxer (:457-473).  The body of
// encode_to_av1 (:749-771) is replaced by: H2D copy -> convert kernel -> RDO kernel (one CTA per
// tile, all tiles of all images of the batch in ONE launch) -> D2H of the block records -> CPU
// symbol packing (thread pool over frames) -> AVIF mux.
//
// No CPU fallback: every encode entry point fails with B200AVIF_ENCODING_ERROR when CUDA is
// unavailable.
#include "../../include/b200avif.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <sched.h>
#include <thread>
#include <vector>

#include "av1_pack.h"
#include "avif_mux.h"
#include "b200_dev.h"
#include "ravif_settings.h"
#include "rdo.h"

struct b200avif_encoder {
  uint8_t quantizer, alpha_quantizer, speed;
  bool premultiplied_alpha;
  int color_model;
  size_t threads;   // 0 = None
  int alpha_color_mode;
  int output_depth; // 0 auto
  std::vector<uint8_t> exif;
  int device;       // -1: current
};

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define ARENA_CHECK(c) do { if ((c).dev.overflow || (c).pin.overflow) return fail(B200AVIF_ENCODING_ERROR, "internal error: work arena too small for this batch"); } while (0)

std::atomic<uint64_t> g_launches{0}, g_h2d{0}, g_d2h{0}, g_tma{0};
std::mutex g_stat_mu;
double g_ms_rdo = 0, g_ms_total = 0, g_ms_pack = 0;

#define CU(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      g_err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;               \
      return B200AVIF_ENCODING_ERROR;                                                            \
    }                                                                                            \
  } while (0)

// ---- a tiny thread pool for symbol packing -------------------------------------------------------
// Host threads this process may use for packing: the CPUs of its affinity mask divided by the processes that share the
// node (LOCAL_WORLD_SIZE under torchrun: one process per GPU), B200AVIF_PACK_THREADS overrides.  Eight ranks on one node do
// not start 8 x all-cores packer threads.  A cgroup CPU quota is deliberately not applied here: packing is a short burst per
// chunk, and under a quota a burst over all the allowed CPUs finishes sooner than the same work on quota-many threads
// (measured: 8192^2 image, 128 CPUs / quota 16: 30 ms against 55 ms).
unsigned packer_threads() {
  if (const char *ev = getenv("B200AVIF_PACK_THREADS")) { const long v = atol(ev); if (v > 0) return (unsigned)v; }
  size_t n = 0;
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = (size_t)CPU_COUNT(&set);
  if (n == 0) n = std::max(1u, std::thread::hardware_concurrency());
  if (const char *ev = getenv("LOCAL_WORLD_SIZE")) { const long v = atol(ev); if (v > 1) n = (n + (size_t)v - 1) / (size_t)v; }
  return (unsigned)std::max<size_t>(n, 1);
}

void parallel_for(size_t n, const std::function<void(size_t)> &fn) {
  const unsigned hw = packer_threads();
  const size_t nt = std::min<size_t>(hw, n);
  if (nt <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (size_t t = 0; t < nt; t++)
    th.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
  for (auto &t : th) t.join();
}

// ---- reusable device / pinned-host arenas (one context per concurrent call) ----------------------
struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  bool host;
  explicit Arena(bool h) : host(h) {}
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    release();
    const size_t want = bytes + bytes / 8;
    cudaError_t e = host ? cudaHostAlloc((void **)&base, want, cudaHostAllocDefault) : cudaMalloc((void **)&base, want);
    if (e != cudaSuccess) { base = nullptr; cap = 0; g_err = std::string("CUDA alloc failed: ") + cudaGetErrorString(e); return 1; }
    cap = want;
    return 0;
  }
  void release() {
    if (base) { if (host) cudaFreeHost(base); else cudaFree(base); }
    base = nullptr; cap = 0;
  }
  void reset() { used = 0; overflow = false; }
  // Bump allocation.  A request past the reserved capacity does NOT hand out memory beyond the arena: it sets
  // `overflow` (checked by the callers before any device work is queued) and returns the arena base.
  bool overflow = false;
  template <typename T> T *take(size_t count) {
    used = (used + 255) & ~(size_t)255;
    if (used + count * sizeof(T) > cap) { overflow = true; return (T *)base; }
    T *p = (T *)(base + used);
    used += count * sizeof(T);
    return p;
  }
};
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct Context {
  int device = 0;
  cudaStream_t stream = nullptr;      // upload, conversion, open-loop analysis, partition decision
  cudaStream_t stream_hi = nullptr;   // high priority: reconstruction chain, in-loop filters, record download
  Arena dev{false}, pin{true};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_analysis = nullptr;  // stages A + B of the current chunk are done
  std::map<int, B200RateTables *> d_rate;   // per base_q_idx
  struct Mark { const char *name; cudaStream_t stream; cudaEvent_t ev; };
  std::vector<Mark> marks;                  // this issue's marks, resolved by finish_rdo()
  std::vector<cudaEvent_t> mark_pool;
  size_t mark_used = 0;
  const uint32_t *stream_total = nullptr;   // last issue_rdo(): total levels gathered (pinned, valid after the streams are drained)
  size_t stream_capacity = 0;
  // Queued work is drained before anything is freed.  Contexts are destroyed by b200avif_shutdown() or when a call
  // drops a context after an error; NOT from a static destructor at process exit (the CUDA runtime may already be gone).
  void drain() { if (stream) cudaStreamSynchronize(stream); if (stream_hi) cudaStreamSynchronize(stream_hi); }
  ~Context() {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
    cudaSetDevice(device);
    drain();
    for (auto &kv : d_rate) cudaFree(kv.second);
    dev.release(); pin.release();
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    for (auto &e : mark_pool) cudaEventDestroy(e);
    if (ev_analysis) cudaEventDestroy(ev_analysis);
    if (stream) cudaStreamDestroy(stream);
    if (stream_hi) cudaStreamDestroy(stream_hi);
    cudaSetDevice(cur);
  }
};
std::mutex g_ctx_mu;
// ---- per-kernel timing (b200avif_set_profiling): CUDA events between the kernels of a stream -----------------------
std::atomic<int> g_profiling{0};
struct StageStat { double ms = 0; uint64_t launches = 0; };
std::map<std::string, StageStat> g_stage;      // guarded by g_stat_mu
thread_local Context *tl_mark_ctx = nullptr;   // the context whose work the calling thread is issuing
std::vector<std::unique_ptr<Context>> g_ctx_free;

int acquire_context(int device, std::unique_ptr<Context> &out) {
  int cur = 0;
  if (device < 0) { if (cudaGetDevice(&cur) != cudaSuccess) return fail(B200AVIF_ENCODING_ERROR, "no CUDA device available (this library has no CPU fallback)"); device = cur; }
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (size_t i = 0; i < g_ctx_free.size(); i++)
      if (g_ctx_free[i]->device == device) { out = std::move(g_ctx_free[i]); g_ctx_free.erase(g_ctx_free.begin() + i); break; }
  }
  CU(cudaSetDevice(device));
  if (!out) {
    std::unique_ptr<Context> c(new Context());      // destroyed (streams / events released) if any step below fails
    c->device = device;
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_lo));
    CU(cudaStreamCreateWithPriority(&c->stream_hi, cudaStreamNonBlocking, prio_hi));
    for (auto &e : c->ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&c->ev_analysis, cudaEventDisableTiming));
    out = std::move(c);
  }
  return 0;
}
// ok == false (the call failed part-way): the context may still have copies / kernels in flight that read its arenas;
// drain it before the next user can reset those arenas.
void release_context(std::unique_ptr<Context> &c, bool ok = true) {
  if (!c) return;
  if (!ok) c->drain();
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  g_ctx_free.insert(g_ctx_free.begin(), std::move(c));      // most recently used first: its arenas are already sized for the caller's work
}

// ---- TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link against libcuda) --------
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = []() -> TmapEncodeFn {
    if (getenv("B200AVIF_NO_TMA")) return nullptr;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return (TmapEncodeFn)p;
  }();
  return fn;
}
// 2-D map of one sample plane (coded width x height, `stride` samples per row), box = box_w x box_h samples, zero fill
// pair8: describe an 8-bit plane as 16-bit elements (half the width, half the box): same bytes, x coordinates halve
bool encode_plane_tmap(CUtensorMap *out, void *base, size_t bps, size_t cw, size_t ch, size_t stride, int box_w, int box_h, bool pair8) {
  TmapEncodeFn fn = tmap_encoder();
  if (!fn || ((uintptr_t)base & 15) || ((stride * bps) & 15)) return false;
  if (pair8 && bps == 1) { bps = 2; cw /= 2; stride /= 2; box_w /= 2; }
  const cuuint64_t dims[2] = {(cuuint64_t)cw, (cuuint64_t)ch};
  const cuuint64_t strides[1] = {(cuuint64_t)(stride * bps)};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1, 1};
  return fn(out, bps == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// RAII lease of a pooled context: handed back drained unless the call marked itself successful.
struct CtxLease {
  std::unique_ptr<Context> c;
  bool ok = false;
  ~CtxLease() { release_context(c, ok); }
};

// ---- one AV1 frame to encode (colour or alpha of an image) ------------------------------------------
struct FrameJob {
  B200Frame f;
  // device
  // sample planes at the native width: uint8_t for 8-bit frames, uint16_t for 10-bit; stride() samples per row
  void *d_src[3] = {nullptr, nullptr, nullptr}, *d_rec[3] = {nullptr, nullptr, nullptr};
  B200Mi *d_mi = nullptr;
  uint16_t *d_eob[3] = {nullptr, nullptr, nullptr};
  int16_t *d_coef[3] = {nullptr, nullptr, nullptr};      // backing store of the compact level streams (dense capacity)
  uint32_t *d_stream_len[3] = {nullptr, nullptr, nullptr};   // per (frame superblock row, tile column) region
  B200BlockDec *d_dec[B200_LEVELS] = {};
  uint32_t *d_satd[B200_LEVELS] = {};
  uint32_t *d_list[B200_LEVELS] = {}, *d_list_n = nullptr;      // survivor lists of the pruned analysis levels
  int8_t *d_cdef = nullptr;
  void *d_fin[3] = {nullptr, nullptr, nullptr};   // CDEF output: input of the loop-restoration search (and the test seam)
  B200LrUnit *d_lr[3] = {nullptr, nullptr, nullptr}, *h_lr[3] = {nullptr, nullptr, nullptr};
  void *d_lr_out[3] = {nullptr, nullptr, nullptr};   // restored planes: test seam only
  void *d_dbrows[3] = {nullptr, nullptr, nullptr};   // deblocked rows 54..57 of every superblock row (loop restoration input)
  void *d_deb_full[3] = {nullptr, nullptr, nullptr}; // whole deblocked planes: test seam only
  int8_t *h_cdef = nullptr;
  unsigned long long *d_lf_sse = nullptr;      // deblocking level search: [3][B200_LF_CANDIDATES] squared errors
  int32_t *d_lf_out = nullptr, *h_lf_out = nullptr;   // the four chosen levels
  // pinned host
  B200Mi *h_mi = nullptr;
  uint16_t *h_eob[3] = {nullptr, nullptr, nullptr};
  const int16_t *h_stream_base = nullptr;                 // gathered level streams of the chunk (pinned)
  const uint32_t *h_stream_off[3] = {nullptr, nullptr, nullptr};   // per plane: offset of every region into h_stream_base
  std::vector<uint8_t> obu;
  int status = 0;
  size_t n_mi() const { return (size_t)f.mi_rows * f.mi_cols; }
  size_t n_px() const { return n_mi() * 16; }
  size_t bps() const { return f.bit_depth > 8 ? 2 : 1; }                         // bytes per sample
  size_t stride() const { return ((size_t)f.mi_cols * 4 + 63) & ~(size_t)63; }   // samples per plane row: 64-sample aligned rows
  size_t plane_bytes() const { return stride() * f.mi_rows * 4 * bps(); }
  size_t dbrows_bytes() const { return stride() * (size_t)f.sb_rows * 4 * bps(); }
  size_t n_coef() const { return (size_t)f.sb_rows * f.sb_cols * B200_SB_COEFS; }
  size_t n_sb() const { return (size_t)f.sb_rows * f.sb_cols; }
  size_t n_blk(int lv) const { return (size_t)(f.mi_rows >> lv) * (f.mi_cols >> lv); }
  size_t n_regions() const { return (size_t)f.sb_rows * f.tile_cols; }
  size_t n_lr() const { return (size_t)b200_lr_units(f.width) * b200_lr_units(f.height); }
  bool uses_lr() const { return f.enable_restoration && (f.lr_type[0] || f.lr_type[1] || f.lr_type[2]); }
};

size_t job_dev_bytes(const FrameJob &j, bool own_src) {
  const int P = j.f.planes;
  size_t b = 0;
  if (own_src) b += P * align256(j.plane_bytes());
  b += P * align256(j.plane_bytes());                 // rec
  b += align256(j.n_mi() * sizeof(B200Mi));
  b += P * align256(j.n_mi() * 2);
  b += P * align256(j.n_coef() * 2);
  b += P * (align256(j.n_regions() * 4) + j.n_regions() * (sizeof(B200StreamRegionH) + 4) + 512);   // stream lengths, gather regions + offsets
  b += align256(j.n_sb());
  for (int lv = 0; lv < B200_LEVELS; lv++) b += align256(j.n_blk(lv) * sizeof(B200BlockDec)) + align256(j.n_blk(lv) * B200_SATD_WORDS * 4) + align256(j.n_blk(lv) * 4) + 768;
  if (j.uses_lr()) b += P * (align256(j.plane_bytes()) + align256(j.dbrows_bytes()) + align256(j.n_lr() * sizeof(B200LrUnit)));
  // the job descriptor and tile list issue_rdo() uploads (one entry per job / tile)
  b += align256(sizeof(B200RdoJob)) + align256((size_t)j.f.tile_rows * j.f.tile_cols * sizeof(B200TileRef)) + 512 + 3 * sizeof(CUtensorMap) + 256;
  return b + 8192;
}
size_t job_pin_bytes(const FrameJob &j) {
  const int P = j.f.planes;
  return align256(j.n_mi() * sizeof(B200Mi)) + P * align256(j.n_mi() * 2) + P * align256(j.n_coef() * 2) /* share of the chunk's stream buffer */ +
         P * (j.n_regions() * 4 + 256) + align256(j.n_sb()) + P * align256(j.n_lr() * sizeof(B200LrUnit)) + 4096;
}
void job_take(FrameJob &j, Arena &dev, Arena &pin, bool own_src) {
  const int P = j.f.planes;
  for (int p = 0; p < P; p++) {
    if (own_src) j.d_src[p] = dev.take<uint8_t>(j.plane_bytes());
    j.d_rec[p] = dev.take<uint8_t>(j.plane_bytes());
    j.d_eob[p] = dev.take<uint16_t>(j.n_mi());
    j.d_coef[p] = dev.take<int16_t>(j.n_coef());
    j.d_stream_len[p] = dev.take<uint32_t>(j.n_regions());
    j.h_eob[p] = pin.take<uint16_t>(j.n_mi());
  }
  j.d_mi = dev.take<B200Mi>(j.n_mi());
  for (int lv = 0; lv < B200_LEVELS; lv++) {
    j.d_dec[lv] = dev.take<B200BlockDec>(j.n_blk(lv));
    j.d_satd[lv] = dev.take<uint32_t>(j.n_blk(lv) * B200_SATD_WORDS);
    j.d_list[lv] = dev.take<uint32_t>(j.n_blk(lv));
  }
  j.d_list_n = dev.take<uint32_t>(B200_LEVELS);
  j.d_cdef = dev.take<int8_t>(j.n_sb());
  j.h_cdef = pin.take<int8_t>(j.n_sb());
  j.d_lf_sse = dev.take<unsigned long long>(3 * B200_LF_CANDIDATES);
  j.d_lf_out = dev.take<int32_t>(4);
  j.h_lf_out = pin.take<int32_t>(4);
  if (j.uses_lr())
    for (int p = 0; p < P; p++) {
      if (!j.d_fin[p]) j.d_fin[p] = dev.take<uint8_t>(j.plane_bytes());
      j.d_dbrows[p] = dev.take<uint8_t>(j.dbrows_bytes());
      j.d_lr[p] = dev.take<B200LrUnit>(j.n_lr());
      j.h_lr[p] = pin.take<B200LrUnit>(j.n_lr());
    }
  j.h_mi = pin.take<B200Mi>(j.n_mi());
}

// Apply ravif's SpeedTweaks (av1encoder.rs:555-605) + rav1e_config (:662-708) to a frame description.
int configure_frame(B200Frame *f, size_t w, size_t h, int depth, bool mono, int quantizer, int speed, size_t threads,
                    int full_range, bool has_cd, int mc) {
  B200SpeedTweaks tw;
  b200_speed_tweaks(speed, quantizer, &tw);
  size_t tiles = b200_tile_target(threads, w, h, (size_t)tw.min_tile_size);
  tiles = std::min<size_t>(std::max<size_t>(tiles, 1), 4096);
  if (b200av1_frame_setup(f, (int)w, (int)h, depth, mono ? 1 : 3, quantizer, (int)tiles) != 0)
    return fail(B200AVIF_INVALID_ARGUMENT, "unsupported frame geometry");
  f->reduced_tx_set = tw.reduced_tx_set;
  f->rdo_tx_decision = tw.rdo_tx_decision;
  f->partition_min = tw.partition_min;
  // The device RDO currently searches square partitions up to 16x16 (ravif reaches 32x32/64x64 only at
  // speed <= 2 with low quantizers, av1encoder.rs:563-570); larger maxima are clamped.
  f->partition_max = std::min(tw.partition_max, 32);      // 64x64 blocks (speed 1) are clamped to 32x32
  f->partition_min = std::min(f->partition_min, f->partition_max);
  f->fine_directional_intra = tw.fine_directional_intra;
  f->enable_intra_edge_filter = 1;     // rav1e's sequence headers always enable the intra edge filter [EXT]
  // early termination of the partition search (records.h split_gate_bits): exhaustive at the slow presets
  f->split_gate_bits[0] = speed <= 2 ? 0 : 60;
  f->split_gate_bits[1] = speed <= 2 ? 0 : 120;
  f->split_gate_bits[2] = speed <= 2 ? 0 : 200;
  if (!tw.cdef) f->enable_cdef = 0;   // ravif: cdef = low_quality && speed <= 9 (av1encoder.rs:590)
  f->psy = 1;                                    // Tune::Psychovisual, always (av1encoder.rs:694)
  f->lf_search = tw.fast_deblock ? 0 : 1;       // av1encoder.rs:586: the closed-form level only at the fast presets
  f->lr_full = tw.sgr_complexity_full ? 1 : 0;  // :573
  f->luma_rdo_modes = tw.complex_prediction_modes ? 4 : 2;      // :572, :628
  f->enable_restoration = tw.lrf ? 1 : 0;     // loop restoration (lrf flag, av1encoder.rs:589): self-guided filter, 64x64 units
  for (int p = 0; p < 3; p++) f->lr_type[p] = (tw.lrf && p < f->planes) ? 1 : 0;
  f->full_range = full_range;
  f->color_description_present = has_cd ? 1 : 0;
  if (has_cd) { f->color_primaries = 1; f->transfer_characteristics = 13; f->matrix_coefficients = mc; }
  return 0;
}

// Rate tables installed from device memory (b200avif_install_rate_tables: the multi-GPU driver broadcasts rank 0's tables
// over NCCL and every rank installs its received copy), per (device, base_q_idx).  Owned by the library.
std::mutex g_installed_mu;
std::map<std::pair<int, int>, B200RateTables *> g_installed_rate;

// Upload (cached) rate tables for this frame's quantiser.
int rate_tables_for(Context &c, int qidx, B200RateTables **out) {
  {
    std::lock_guard<std::mutex> lk(g_installed_mu);
    auto gi = g_installed_rate.find(std::make_pair(c.device, qidx));
    if (gi != g_installed_rate.end()) { *out = gi->second; return 0; }
  }
  auto it = c.d_rate.find(qidx);
  if (it != c.d_rate.end()) { *out = it->second; return 0; }
  B200RateTables host;
  b200av1_build_rate_tables(qidx, &host);
  B200RateTables *d = nullptr;
  CU(cudaMalloc((void **)&d, sizeof(host)));
  // Stream-ordered upload + wait.  A plain cudaMemcpy from pageable memory returns once the data is STAGED: the DMA into
  // `d` is ordered only against the legacy default stream, not against this context's non-blocking streams, so kernels
  // launched right after could read the table before it arrived (seen as rare payload differences when concurrent calls
  // make a context build a table it had not cached yet).
  CU(cudaMemcpyAsync(d, &host, sizeof(host), cudaMemcpyHostToDevice, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  c.d_rate[qidx] = d;
  *out = d;
  return 0;
}

// Issue (asynchronously, on c.stream) the RDO stages + in-loop filters for a set of jobs whose device buffers are
// ready and sources filled, and the D2H copies of the block records.  finish_rdo() waits for it.
// Stages A + B are queued on c.stream (behind `after_analysis_of`, the previous chunk's ev_analysis, so that the analyses
// of successive chunks run back to back); stage C, the filters and the download go to the high-priority c.stream_hi and
// overlap the next chunk's analysis.
int issue_rdo(Context &c, std::vector<FrameJob *> &jobs, Arena &dev, Arena &pin, bool fetch_rec,
              std::vector<std::vector<uint8_t>> *rec_out, std::vector<std::vector<uint8_t>> *pre_filter_rec = nullptr,
              cudaEvent_t after_analysis_of = nullptr) {
  struct MarkScope { MarkScope(Context *c) { tl_mark_ctx = c; c->marks.clear(); c->mark_used = 0; } ~MarkScope() { tl_mark_ctx = nullptr; } } mark_scope(&c);
  std::vector<B200RdoJob> hj(jobs.size());
  std::vector<B200TileRef> tiles;
  for (size_t i = 0; i < jobs.size(); i++) {
    FrameJob &j = *jobs[i];
    B200RdoJob &r = hj[i];
    memset(&r, 0, sizeof(r));
    r.frame = j.f;
    r.stride = (int32_t)j.stride();
    B200RateTables *rt;
    if (int e = rate_tables_for(c, j.f.base_q_idx, &rt)) return e;
    r.rate = rt;
    for (int p = 0; p < 3; p++) {
      const int q = p < j.f.planes ? p : 0;
      r.src[p] = j.d_src[q]; r.rec[p] = j.d_rec[q]; r.eob[p] = j.d_eob[q]; r.coef[p] = j.d_coef[q]; r.stream_len[p] = j.d_stream_len[q];
    }
    r.mi = j.d_mi;
    for (int lv = 0; lv < B200_LEVELS; lv++) { r.dec[lv] = j.d_dec[lv]; r.satd[lv] = j.d_satd[lv]; r.list[lv] = j.d_list[lv]; }
    r.list_n = j.d_list_n;
    r.cdef_idx = j.d_cdef;
    r.lf_sse = j.d_lf_sse; r.lf_out = j.d_lf_out;
    for (int p = 0; p < 3; p++) { r.fin[p] = j.d_fin[p < j.f.planes ? p : 0]; r.dbrows[p] = p < j.f.planes ? j.d_dbrows[p] : nullptr; r.deb_full[p] = p < j.f.planes ? j.d_deb_full[p] : nullptr; }
    for (int p = 0; p < 3; p++) { r.lr[p] = p < j.f.planes ? j.d_lr[p] : nullptr; r.lr_out[p] = p < j.f.planes ? j.d_lr_out[p] : nullptr; }
    for (int tr = 0; tr < j.f.tile_rows; tr++)
      for (int tc = 0; tc < j.f.tile_cols; tc++) tiles.push_back(B200TileRef{(int32_t)i, (int16_t)tr, (int16_t)tc});
  }
  B200RdoJob *d_jobs = dev.take<B200RdoJob>(hj.size());
  B200TileRef *d_tiles = dev.take<B200TileRef>(tiles.size());
  if (dev.overflow || pin.overflow) return fail(B200AVIF_ENCODING_ERROR, "internal error: work arena too small for this batch");
  CU(cudaMemcpyAsync(d_jobs, hj.data(), hj.size() * sizeof(B200RdoJob), cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(B200TileRef), cudaMemcpyHostToDevice, c.stream));
  if (after_analysis_of) CU(cudaStreamWaitEvent(c.stream, after_analysis_of, 0));
  CU(cudaEventRecord(c.ev[1], c.stream));
  int max_blocks4 = 0, top_level = 2;
  for (FrameJob *jp : jobs) { max_blocks4 = std::max(max_blocks4, (int)jp->n_mi()); if (jp->f.partition_max >= 32) top_level = 3; }
  int max_regions = 0;       // root regions of the partition decision: every job decides with ITS root size (16 or 32)
  for (FrameJob *jp : jobs) {
    const int rn4 = jp->f.partition_max >= 32 ? 8 : 4;
    max_regions = std::max(max_regions, ((jp->f.mi_cols + rn4 - 1) / rn4) * ((jp->f.mi_rows + rn4 - 1) / rn4));
  }
  // superblock-row slots of the reconstruction chain: as many as keep every tile of the launch resident at once
  int recon_rows = 4;
  {
    int dev_id = 0, sms = 148, smem_sm = 227 * 1024;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev_id);
    int max_tile_rows = 1;
    for (FrameJob *jp : jobs)
      for (int tr = 0; tr < jp->f.tile_rows; tr++) max_tile_rows = std::max(max_tile_rows, jp->f.tile_row_start_sb[tr + 1] - jp->f.tile_row_start_sb[tr]);
    while (recon_rows > 1) {
      const size_t need = b200_recon_smem_bytes(jobs[0]->f.bit_depth > 8, top_level, recon_rows) + 1024;
      const size_t per_sm = need ? (size_t)smem_sm / need : 0;
      if (recon_rows <= max_tile_rows && per_sm * (size_t)sms >= tiles.size()) break;
      recon_rows--;
    }
    if (const char *ev = getenv("B200AVIF_RECON_ROWS")) {      // A/B and test knob; never more slots than one SM's shared memory holds
      const int v = atoi(ev);
      if (v >= 1 && v <= 4) recon_rows = v;
      while (recon_rows > 1 && b200_recon_smem_bytes(jobs[0]->f.bit_depth > 8, top_level, recon_rows) + 1024 > (size_t)smem_sm) recon_rows--;
    }
  }
  int n_launched = 0;
  const int le = b200_rdo_launch(d_jobs, (int)jobs.size(), d_tiles, (int)tiles.size(), max_blocks4, max_regions, jobs[0]->f.bit_depth > 8,
                                 top_level, recon_rows, c.stream, c.stream_hi, c.ev_analysis, &n_launched);   // one bit depth per call
  if (le != 0) return fail(B200AVIF_ENCODING_ERROR, std::string("RDO kernel launch failed: ") + cudaGetErrorString((cudaError_t)le));
  g_launches += (uint64_t)n_launched;   // counted by the launcher: (SATD + analysis) per level + partition decision + reconstruction
  CU(cudaEventRecord(c.ev[2], c.stream_hi));
  if (pre_filter_rec) {   // test seam: reconstruction before the in-loop filters
    pre_filter_rec->clear();
    for (FrameJob *jp : jobs)
      for (int p = 0; p < jp->f.planes; p++) {
        pre_filter_rec->emplace_back(jp->plane_bytes());
        CU(cudaMemcpyAsync(pre_filter_rec->back().data(), jp->d_rec[p], jp->plane_bytes(), cudaMemcpyDeviceToHost, c.stream_hi));
      }
    CU(cudaStreamSynchronize(c.stream_hi));
  }
  {
    int max_sb = 0;
    for (FrameJob *jp : jobs) max_sb = std::max(max_sb, (int)jp->n_sb());
    int with_lr = 0, with_lf = 0;
    for (FrameJob *jp : jobs) { with_lr |= jp->uses_lr(); with_lf |= jp->f.lf_search != 0; }
    if (with_lf)
      for (FrameJob *jp : jobs) CU(cudaMemsetAsync(jp->d_lf_sse, 0, sizeof(unsigned long long) * 3 * B200_LF_CANDIDATES, c.stream_hi));
    // tensor maps of the reconstruction planes for the TMA loads of the fused filter kernel (one per job and plane)
    const void *d_tmaps = nullptr;
    static const bool tma_pair8 = getenv("B200AVIF_TMA_PAIR8") != nullptr;
    {
      int bw = 0, bh = 0;
      b200_filters_tmap_box(jobs[0]->f.bit_depth > 8, &bw, &bh);
      std::vector<CUtensorMap> maps(jobs.size() * 3);
      bool ok = tmap_encoder() != nullptr;
      if (jobs[0]->f.bit_depth == 8 && getenv("B200AVIF_TMA_SKIP_U8")) ok = false;
      for (size_t i = 0; ok && i < jobs.size(); i++)
        for (int p = 0; ok && p < 3; p++) {
          FrameJob &j = *jobs[i];
          const int q = p < j.f.planes ? p : 0;
          ok = encode_plane_tmap(&maps[i * 3 + p], j.d_rec[q], j.bps(), (size_t)j.f.mi_cols * 4, (size_t)j.f.mi_rows * 4, j.stride(), bw, bh, tma_pair8);
        }
      if (ok) {
        CUtensorMap *d = dev.take<CUtensorMap>(maps.size());
        if (dev.overflow) return fail(B200AVIF_ENCODING_ERROR, "internal error: work arena too small for this batch");
        CU(cudaMemcpyAsync(d, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, c.stream_hi));
        d_tmaps = d;
        g_tma++;
      }
    }
    b200_mark(nullptr, c.stream_hi);
    const int fe = b200_filters_launch(d_jobs, (int)jobs.size(), max_sb, with_lr, with_lf, jobs[0]->f.bit_depth > 8, d_tmaps, tma_pair8 ? 1 : 0, c.stream_hi, &n_launched);
    if (fe != 0) return fail(B200AVIF_ENCODING_ERROR, std::string("filter kernel launch failed: ") + cudaGetErrorString((cudaError_t)fe));
    g_launches += (uint64_t)n_launched;
  }
  uint64_t d2h = 0;
  for (FrameJob *jp : jobs) {
    FrameJob &j = *jp;
    CU(cudaMemcpyAsync(j.h_mi, j.d_mi, j.n_mi() * sizeof(B200Mi), cudaMemcpyDeviceToHost, c.stream_hi));
    CU(cudaMemcpyAsync(j.h_cdef, j.d_cdef, j.n_sb(), cudaMemcpyDeviceToHost, c.stream_hi));
    if (j.f.lf_search) CU(cudaMemcpyAsync(j.h_lf_out, j.d_lf_out, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream_hi));
    d2h += j.n_mi() * sizeof(B200Mi) + j.n_sb();
    if (j.uses_lr())
      for (int p = 0; p < j.f.planes; p++) {
        CU(cudaMemcpyAsync(j.h_lr[p], j.d_lr[p], j.n_lr() * sizeof(B200LrUnit), cudaMemcpyDeviceToHost, c.stream_hi));
        d2h += j.n_lr() * sizeof(B200LrUnit);
      }
    for (int p = 0; p < j.f.planes; p++) {
      CU(cudaMemcpyAsync(j.h_eob[p], j.d_eob[p], j.n_mi() * 2, cudaMemcpyDeviceToHost, c.stream_hi));
      d2h += j.n_mi() * 2;
    }
  }
  {
    // Quantised levels: the reconstruction chain wrote compact streams (eob levels per transform block) at the start of
    // every (frame superblock row, tile column) region of the coefficient planes.  One scan + one gather kernel pack all
    // regions of the chunk back to back straight into pinned host memory (the used bytes only cross PCIe), the packer
    // gets the offsets.
    std::vector<B200StreamRegionH> regs;
    size_t capacity = 0;
    for (FrameJob *jp : jobs)
      for (int p = 0; p < jp->f.planes; p++) {
        capacity += jp->n_coef();
        for (int sbr = 0; sbr < jp->f.sb_rows; sbr++)
          for (int tc = 0; tc < jp->f.tile_cols; tc++)
            regs.push_back(B200StreamRegionH{jp->d_coef[p] + ((size_t)sbr * jp->f.sb_cols + jp->f.tile_col_start_sb[tc]) * B200_SB_COEFS,
                                             jp->d_stream_len[p] + (size_t)sbr * jp->f.tile_cols + tc});
      }
    capacity = std::min<size_t>(capacity, 0xfffffff0u);
    B200StreamRegionH *d_regs = dev.take<B200StreamRegionH>(regs.size());
    uint32_t *d_off = dev.take<uint32_t>(regs.size() + 1);
    uint32_t *h_off = pin.take<uint32_t>(regs.size() + 1);
    int16_t *h_stream = pin.take<int16_t>(capacity);
    if (dev.overflow || pin.overflow) return fail(B200AVIF_ENCODING_ERROR, "internal error: work arena too small for this batch");
    CU(cudaMemcpyAsync(d_regs, regs.data(), regs.size() * sizeof(B200StreamRegionH), cudaMemcpyHostToDevice, c.stream_hi));
    const int ge = b200_stream_gather_launch(d_regs, (int)regs.size(), d_off, h_stream, (uint32_t)capacity, c.stream_hi);
    if (ge != 0) return fail(B200AVIF_ENCODING_ERROR, std::string("stream gather launch failed: ") + cudaGetErrorString((cudaError_t)ge));
    g_launches += 2;
    CU(cudaMemcpyAsync(h_off, d_off, (regs.size() + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream_hi));
    d2h += (regs.size() + 1) * sizeof(uint32_t);
    size_t k = 0;
    for (FrameJob *jp : jobs)
      for (int p = 0; p < jp->f.planes; p++) { jp->h_stream_base = h_stream; jp->h_stream_off[p] = h_off + k; k += jp->n_regions(); }
    c.stream_total = h_off + regs.size();
    c.stream_capacity = capacity;
  }
  if (fetch_rec && rec_out) {
    rec_out->clear();
    for (FrameJob *jp : jobs)
      for (int p = 0; p < jp->f.planes; p++) {
        rec_out->emplace_back(jp->plane_bytes());
        CU(cudaMemcpyAsync(rec_out->back().data(), jp->d_deb_full[p] ? jp->d_deb_full[p] : jp->d_rec[p], jp->plane_bytes(), cudaMemcpyDeviceToHost, c.stream_hi));
      }
  }
  CU(cudaEventRecord(c.ev[3], c.stream_hi));
  g_d2h += d2h;
  return 0;
}

// Wait for the device work issued by issue_rdo() on this context and account its CUDA-event times.
int finish_rdo(Context &c) {
  CU(cudaStreamSynchronize(c.stream_hi));
  CU(cudaStreamSynchronize(c.stream));
  if (c.stream_total) {
    if (*c.stream_total > c.stream_capacity) return fail(B200AVIF_ENCODING_ERROR, "internal error: level stream larger than its buffer");
    g_d2h += (uint64_t)*c.stream_total * 2;
    c.stream_total = nullptr;
  }
  if (!c.marks.empty()) {       // time between consecutive marks of a stream -> the kernel launched in between
    std::map<cudaStream_t, cudaEvent_t> last;
    std::lock_guard<std::mutex> lk(g_stat_mu);
    for (const Context::Mark &m : c.marks) {
      auto it = last.find(m.stream);
      if (m.name && it != last.end()) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, it->second, m.ev) == cudaSuccess) { StageStat &st = g_stage[m.name]; st.ms += ms; st.launches++; }
      }
      last[m.stream] = m.ev;
    }
    c.marks.clear();
  }
  float ms_rdo = 0, ms_tot = 0;
  cudaEventElapsedTime(&ms_rdo, c.ev[1], c.ev[2]);
  cudaEventElapsedTime(&ms_tot, c.ev[0], c.ev[3]);
  std::lock_guard<std::mutex> lk(g_stat_mu);
  g_ms_rdo += ms_rdo; g_ms_total += ms_tot;
  return 0;
}

int run_rdo_and_fetch(Context &c, std::vector<FrameJob *> &jobs, Arena &dev, Arena &pin, bool fetch_rec,
                      std::vector<std::vector<uint8_t>> *rec_out, std::vector<std::vector<uint8_t>> *pre_filter_rec = nullptr) {
  if (int e = issue_rdo(c, jobs, dev, pin, fetch_rec, rec_out, pre_filter_rec)) return e;
  return finish_rdo(c);
}

void pack_jobs(std::vector<FrameJob *> &jobs) {
  const auto t0 = std::chrono::steady_clock::now();
  {   // few frames (one large image): spread each frame's tiles over the idle host threads
    const unsigned hw = packer_threads();
    b200av1_pack_set_tile_threads(jobs.size() * 2 <= hw ? (int)(hw / std::max<size_t>(jobs.size(), 1)) : 1);
  }
  parallel_for(jobs.size(), [&](size_t i) {
    FrameJob &j = *jobs[i];
    if (j.f.lf_search) for (int k = 0; k < 4; k++) j.f.lf_level[k] = j.h_lf_out[k];      // levels chosen on the device -> frame header
    B200Records rec;
    memset(&rec, 0, sizeof(rec));
    rec.mi = j.h_mi;
    rec.cdef_idx = j.h_cdef;
    for (int p = 0; p < j.f.planes; p++) rec.lr[p] = j.h_lr[p];
    for (int p = 0; p < j.f.planes; p++) { rec.eob[p] = j.h_eob[p]; rec.stream_off[p] = j.h_stream_off[p]; }
    rec.stream_base = j.h_stream_base;
    uint8_t *out = nullptr;
    size_t len = 0;
    j.status = b200av1_pack_frame(&j.f, &rec, &out, &len);
    if (j.status == 0) { j.obu.assign(out, out + len); free(out); }
  });
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::lock_guard<std::mutex> lk(g_stat_mu);
  g_ms_pack += ms;
}

struct ImageWork {
  // input
  const uint8_t *pixels = nullptr; bool on_device = false; int channels = 4;
  size_t w = 0, h = 0, stride = 0;
  // raw planes input (encode_raw_planes_*)
  bool raw = false; std::vector<uint16_t> raw_planes[3], raw_alpha; bool raw_has_alpha = false;
  int raw_range = 1, raw_mc = 6;
  // work
  uint8_t *d_in = nullptr; int *d_flag = nullptr;
  int depth = 10;
  bool use_alpha = false;
  FrameJob color, alpha;
  int status = 0;
};

int encode_images(const b200avif_encoder &enc, std::vector<ImageWork> &imgs, b200avif_image *outs, int *status) {
  // every output is defined on every return path: a caller may b200avif_image_free() a failed call's images
  for (size_t i = 0; i < imgs.size(); i++) { memset(&outs[i], 0, sizeof(outs[i])); if (status) status[i] = B200AVIF_ENCODING_ERROR; }
  if (b200avif_device_count() <= 0)
    return fail(B200AVIF_ENCODING_ERROR, "no CUDA device available (this library has no CPU fallback)");
  constexpr size_t NC = 3;            // contexts: two chunks in flight on the device + one being packed / refilled
  CtxLease lease[NC];                 // destroyed last: after the packer threads were joined (Joiner below)
  if (int e = acquire_context(enc.device, lease[0].c)) return e;
  std::unique_ptr<Context> &ctx = lease[0].c;
  const int depth = enc.output_depth == 8 ? 8 : 10;
  const int mc = enc.color_model == B200AVIF_COLOR_YCBCR ? B200AVIF_MC_BT601 : B200AVIF_MC_IDENTITY;

  for (ImageWork &im : imgs) im.depth = depth;
  // ---- frame descriptions
  for (ImageWork &im : imgs) {
    const int cmc = im.raw ? im.raw_mc : mc;
    const int range = im.raw ? im.raw_range : 1;
    if (int e = configure_frame(&im.color.f, im.w, im.h, depth, false, enc.quantizer, enc.speed, enc.threads, range, true, cmc)) { im.status = e; continue; }
    if (int e = configure_frame(&im.alpha.f, im.w, im.h, depth, true, enc.alpha_quantizer, enc.speed, enc.threads, 1, false, 2)) { im.status = e; continue; }
  }
  // ---- software pipeline over chunks of the batch on two contexts (stream + arenas each):
  //      upload/convert of chunk k+1 overlaps the RDO of chunk k, symbol packing (CPU threads) of chunk k overlaps
  //      the device work of chunk k+1.
  const size_t n_img = imgs.size();
  size_t n_chunks = n_img >= 16 ? 4 : (n_img >= 4 ? 2 : 1);
  if (const char *ev = getenv("B200AVIF_CHUNKS")) { const long v = atol(ev); if (v >= 1) n_chunks = std::min<size_t>((size_t)v, n_img); }
  for (size_t k = 1; k < std::min(NC, n_chunks); k++) { if (int e = acquire_context(enc.device, lease[k].c)) return e; }
  Context *cx[NC];
  for (size_t k = 0; k < NC; k++) cx[k] = lease[k].c ? lease[k].c.get() : ctx.get();
  struct Chunk { size_t a, b; int *h_flags = nullptr; std::vector<FrameJob *> jobs; std::thread packer; bool issued = false; };
  std::vector<Chunk> chunks(n_chunks);
  for (size_t k = 0; k < n_chunks; k++) { chunks[k].a = k * n_img / n_chunks; chunks[k].b = (k + 1) * n_img / n_chunks; }
  auto join_all = [&]() { for (Chunk &ch : chunks) if (ch.packer.joinable()) ch.packer.join(); };
  struct Joiner { std::function<void()> f; ~Joiner() { f(); } } joiner{join_all};

  auto phase1 = [&](Context &c, Chunk &ch) -> int {       // upload + colour conversion + alpha flags (async)
    size_t dev_bytes = 1 << 20, pin_bytes = 1 << 20;
    for (size_t i = ch.a; i < ch.b; i++) {
      ImageWork &im = imgs[i];
      if (im.status) continue;
      dev_bytes += job_dev_bytes(im.color, true) + job_dev_bytes(im.alpha, true);
      pin_bytes += job_pin_bytes(im.color) + job_pin_bytes(im.alpha);
      const size_t in_bytes = im.raw ? (im.w * im.h * 2 * 4) : (im.stride * im.h * im.channels);
      dev_bytes += align256(in_bytes) + 512;
      if (!im.raw && im.channels == 4) dev_bytes += 2 * align256(im.w * im.h * 4) + 512;    // alpha pre-clean scratch
      if (!im.on_device) pin_bytes += align256(in_bytes) + 512;
    }
    if (c.dev.reserve(dev_bytes) || c.pin.reserve(pin_bytes)) return B200AVIF_ENCODING_ERROR;
    c.dev.reset(); c.pin.reset();
    CU(cudaEventRecord(c.ev[0], c.stream));
    uint64_t h2d = 0;
    const size_t cnt = ch.b - ch.a;
    ch.h_flags = c.pin.take<int>(cnt);
    int *d_flags = c.dev.take<int>(cnt);
    CU(cudaMemsetAsync(d_flags, 0, cnt * sizeof(int), c.stream));
    // pageable caller memory goes through the pinned arena: every such image of the chunk is staged by ONE parallel copy
    // (1 MB pieces over up to 16 host threads) before the uploads are issued; page-locked caller memory is copied straight
    std::vector<uint8_t *> staged(cnt, nullptr);
    {
      struct Piece { uint8_t *dst; const uint8_t *src; size_t n; };
      std::vector<Piece> pieces;
      for (size_t i = ch.a; i < ch.b; i++) {
        ImageWork &im = imgs[i];
        if (im.status || im.raw || im.on_device) continue;
        const size_t bytes = (im.stride * (im.h - 1) + im.w) * im.channels;
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, im.pixels) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) cudaGetLastError();
        if (pinned) continue;
        uint8_t *stage = c.pin.take<uint8_t>(bytes);
        ARENA_CHECK(c);
        staged[i - ch.a] = stage;
        for (size_t o = 0; o < bytes; o += (1u << 20)) pieces.push_back({stage + o, im.pixels + o, std::min<size_t>(1u << 20, bytes - o)});
      }
      const size_t nt = std::min<size_t>(std::min<size_t>(16, packer_threads()), pieces.size());
      if (nt <= 1) { for (const Piece &pc : pieces) memcpy(pc.dst, pc.src, pc.n); }
      else parallel_for(nt, [&](size_t t) { for (size_t k = t; k < pieces.size(); k += nt) memcpy(pieces[k].dst, pieces[k].src, pieces[k].n); });
    }
    for (size_t i = ch.a; i < ch.b; i++) {
      ImageWork &im = imgs[i];
      if (im.status) continue;
      job_take(im.color, c.dev, c.pin, true);
      job_take(im.alpha, c.dev, c.pin, true);
      ARENA_CHECK(c);
      const int cw = im.color.f.mi_cols * 4, chh = im.color.f.mi_rows * 4;
      if (im.raw) {
        for (int p = 0; p < 4; p++) {
          if (p == 3 && !im.raw_has_alpha) break;
          const std::vector<uint16_t> &srcv = p < 3 ? im.raw_planes[p] : im.raw_alpha;
          uint16_t *stage = c.pin.take<uint16_t>(im.w * im.h);
          uint16_t *d = c.dev.take<uint16_t>(im.w * im.h);
          ARENA_CHECK(c);
          memcpy(stage, srcv.data(), im.w * im.h * 2);
          CU(cudaMemcpyAsync(d, stage, im.w * im.h * 2, cudaMemcpyHostToDevice, c.stream));
          if (int e = b200_pad_plane_launch(d, (int)im.w, (int)im.h, (int)im.w, p < 3 ? im.color.d_src[p] : im.alpha.d_src[0], cw, chh, (int)im.color.stride(), depth, c.stream))
            return fail(B200AVIF_ENCODING_ERROR, cudaGetErrorString((cudaError_t)e));
          g_launches++; h2d += im.w * im.h * 2;
        }
        continue;
      }
      const uint8_t *d_in;
      if (im.on_device) d_in = im.pixels;
      else {
        const size_t bytes = (im.stride * (im.h - 1) + im.w) * im.channels;
        const uint8_t *from = staged[i - ch.a] ? staged[i - ch.a] : im.pixels;      // staged above unless page-locked
        uint8_t *d = c.dev.take<uint8_t>(bytes + 16);
        ARENA_CHECK(c);
        CU(cudaMemcpyAsync(d, from, bytes, cudaMemcpyHostToDevice, c.stream));
        h2d += bytes;
        d_in = d;
      }
      size_t in_stride = im.stride;
      if (im.channels == 4 && enc.alpha_color_mode != B200AVIF_ALPHA_UNASSOCIATED_DIRTY) {
        // alpha pre-clean on the device (av1encoder.rs:277-299, dirtyalpha.rs): the converter reads the cleaned copy
        const int mode = enc.alpha_color_mode == B200AVIF_ALPHA_PREMULTIPLIED ? 2 : 1;
        uint8_t *d_tmp = c.dev.take<uint8_t>(im.w * im.h * 4), *d_clean = c.dev.take<uint8_t>(im.w * im.h * 4);
        unsigned long long *d_acc = c.dev.take<unsigned long long>(4);
        ARENA_CHECK(c);
        if (int e = b200_alpha_clean_launch(d_in, (int)im.w, (int)im.h, (int)im.stride, mode, d_tmp, d_clean, d_acc, c.stream))
          return fail(B200AVIF_ENCODING_ERROR, cudaGetErrorString((cudaError_t)e));
        g_launches += mode == 2 ? 1 : 3;
        d_in = d_clean; in_stride = im.w;
      }
      if (int e = b200_convert_launch(d_in, im.channels, (int)im.w, (int)im.h, (int)in_stride, cw, chh, (int)im.color.stride(), depth, enc.color_model,
                                      im.color.d_src[0], im.color.d_src[1], im.color.d_src[2],
                                      im.channels == 4 ? im.alpha.d_src[0] : nullptr, im.channels == 4 ? d_flags + (i - ch.a) : nullptr, c.stream))
        return fail(B200AVIF_ENCODING_ERROR, cudaGetErrorString((cudaError_t)e));
      g_launches++;
    }
    g_h2d += h2d;
    CU(cudaMemcpyAsync(ch.h_flags, d_flags, cnt * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ch.issued = true;
    return 0;
  };
  auto phase2 = [&](Context &c, Chunk &ch, cudaEvent_t prev_analysis) -> int {       // needs the alpha flags: RDO + filters + record D2H (async)
    CU(cudaStreamSynchronize(c.stream));
    for (size_t i = ch.a; i < ch.b; i++) {
      ImageWork &im = imgs[i];
      if (im.status) continue;
      im.use_alpha = im.raw ? im.raw_has_alpha : (im.channels == 4 && ch.h_flags[i - ch.a] != 0);
      ch.jobs.push_back(&im.color);
      if (im.use_alpha) ch.jobs.push_back(&im.alpha);
    }
    if (ch.jobs.empty()) return 0;
    return issue_rdo(c, ch.jobs, c.dev, c.pin, false, nullptr, nullptr, prev_analysis);
  };
  auto phase3 = [&](Context &c, Chunk &ch) -> int {       // wait for the device, hand the records to the packer thread
    if (ch.jobs.empty()) return 0;
    if (int e = finish_rdo(c)) return e;
    ch.packer = std::thread([&ch]() { pack_jobs(ch.jobs); });
    return 0;
  };
  cudaEvent_t prev_analysis = nullptr;
  for (size_t k = 0; k < n_chunks; k++) {
    Context &c = *cx[k % NC];
    if (k >= NC && chunks[k - NC].packer.joinable()) chunks[k - NC].packer.join();   // its pinned arena is about to be reused
    if (int e = phase1(c, chunks[k])) return e;
    if (int e = phase2(c, chunks[k], prev_analysis)) return e;      // queue this chunk's device work before waiting for older ones
    if (!chunks[k].jobs.empty()) prev_analysis = c.ev_analysis;
    if (k + 1 >= NC) { const size_t d = k + 1 - NC; if (int e = phase3(*cx[d % NC], chunks[d])) return e; }
  }
  for (size_t d = n_chunks >= NC ? n_chunks - NC + 1 : 0; d < n_chunks; d++)
    if (int e = phase3(*cx[d % NC], chunks[d])) return e;
  join_all();
  // ---- mux
  const auto t0 = std::chrono::steady_clock::now();
  int first_err = 0;
  for (size_t k = 0; k < NC; k++) lease[k].ok = true;      // all device work of this call has been waited for
  for (size_t i = 0; i < imgs.size(); i++) {
    ImageWork &im = imgs[i];
    if (!im.status && (im.color.status || (im.use_alpha && im.alpha.status))) im.status = B200AVIF_ENCODING_ERROR;
    if (!im.status) {
      B200AvifMuxInput mi;
      memset(&mi, 0, sizeof(mi));
      mi.color = im.color.obu.data(); mi.color_len = im.color.obu.size();
      if (im.use_alpha) { mi.alpha = im.alpha.obu.data(); mi.alpha_len = im.alpha.obu.size(); }
      if (!enc.exif.empty()) { mi.exif = enc.exif.data(); mi.exif_len = enc.exif.size(); }
      mi.width = (uint32_t)im.w; mi.height = (uint32_t)im.h; mi.depth = depth;
      mi.matrix_coefficients = im.raw ? im.raw_mc : mc;
      mi.premultiplied_alpha = enc.premultiplied_alpha;
      mi.color_level_idx = b200av1_seq_level_idx(&im.color.f);
      mi.alpha_level_idx = b200av1_seq_level_idx(&im.alpha.f);
      if (b200avif_mux(&mi, &outs[i].avif_file, &outs[i].avif_len) != 0) im.status = B200AVIF_ENCODING_ERROR;
      outs[i].color_byte_size = im.color.obu.size();
      outs[i].alpha_byte_size = im.use_alpha ? im.alpha.obu.size() : 0;
    }
    if (status) status[i] = im.status;
    if (im.status && !first_err) first_err = im.status;
  }
  {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> lk(g_stat_mu);
    g_ms_pack += ms;
  }
  return first_err;
}

bool valid_mc(int mc) {
  return mc == 0 || mc == 1 || mc == 2 || mc == 6 || mc == 8 || mc == 9 || mc == 10;
}

template <typename P>
int encode_raw(const b200avif_encoder *enc, size_t w, size_t h, const P *planes, size_t n_px, const P *alpha, size_t n_alpha,
               int range, int mc, int depth, b200avif_image *out) {
  if (!enc || !out || !planes || w == 0 || h == 0) return fail(B200AVIF_INVALID_ARGUMENT, "null/zero argument");
  if (!valid_mc(mc)) return fail(B200AVIF_UNSUPPORTED, "matrix coefficients");   // av1encoder.rs:468
  if (n_px < w * h || (alpha && n_alpha < w * h)) return fail(B200AVIF_TOO_FEW_PIXELS, "Provided buffer is smaller than width * height");
  std::vector<ImageWork> imgs(1);
  ImageWork &im = imgs[0];
  im.raw = true; im.w = w; im.h = h; im.raw_range = range; im.raw_mc = mc;
  for (int p = 0; p < 3; p++) {
    im.raw_planes[p].resize(w * h);
    for (size_t i = 0; i < w * h; i++) im.raw_planes[p][i] = planes[i * 3 + p];
  }
  if (alpha) { im.raw_has_alpha = true; im.raw_alpha.assign(alpha, alpha + w * h); }
  b200avif_encoder e2 = *enc;
  e2.output_depth = depth;
  return encode_images(e2, imgs, out, nullptr);
}

}  // namespace

// ================================================================================================
extern "C" {

b200avif_encoder *b200avif_encoder_new(void) {
  b200avif_encoder *e = new b200avif_encoder();
  e->quantizer = b200_quality_to_quantizer(80.f);          // av1encoder.rs:88-102
  e->alpha_quantizer = b200_quality_to_quantizer(80.f);
  e->speed = 5;
  e->output_depth = 0;
  e->premultiplied_alpha = false;
  e->color_model = B200AVIF_COLOR_YCBCR;
  e->threads = 0;
  e->alpha_color_mode = B200AVIF_ALPHA_UNASSOCIATED_CLEAN;
  e->device = -1;
  return e;
}
b200avif_encoder *b200avif_encoder_clone(const b200avif_encoder *enc) { return enc ? new b200avif_encoder(*enc) : nullptr; }
void b200avif_encoder_free(b200avif_encoder *enc) { delete enc; }

int b200avif_encoder_with_quality(b200avif_encoder *enc, float q) {
  if (!enc || !(q >= 1.f && q <= 100.f)) return fail(B200AVIF_INVALID_ARGUMENT, "quality must be in 1..=100");
  enc->quantizer = b200_quality_to_quantizer(q);
  return 0;
}
int b200avif_encoder_with_alpha_quality(b200avif_encoder *enc, float q) {
  if (!enc || !(q >= 1.f && q <= 100.f)) return fail(B200AVIF_INVALID_ARGUMENT, "alpha quality must be in 1..=100");
  enc->alpha_quantizer = b200_quality_to_quantizer(q);
  return 0;
}
// The reference struct stores the QUANTIZERS (av1encoder.rs:67-71: quantizer, alpha_quantizer: u8), not the qualities:
// a binding that mirrors the struct field by field uses these two.
int b200avif_encoder_with_quantizer(b200avif_encoder *enc, uint8_t quantizer) {
  if (!enc) return fail(B200AVIF_INVALID_ARGUMENT, "null encoder");
  enc->quantizer = quantizer;
  return 0;
}
int b200avif_encoder_with_alpha_quantizer(b200avif_encoder *enc, uint8_t quantizer) {
  if (!enc) return fail(B200AVIF_INVALID_ARGUMENT, "null encoder");
  enc->alpha_quantizer = quantizer;
  return 0;
}
int b200avif_encoder_with_speed(b200avif_encoder *enc, uint8_t s) {
  if (!enc || s < 1 || s > 10) return fail(B200AVIF_INVALID_ARGUMENT, "speed must be in 1..=10");
  enc->speed = s;
  return 0;
}
int b200avif_encoder_with_bit_depth(b200avif_encoder *enc, int d) {
  if (!enc || (d != 0 && d != 8 && d != 10)) return fail(B200AVIF_INVALID_ARGUMENT, "bit depth must be 0 (auto), 8 or 10");
  enc->output_depth = d;
  return 0;
}
int b200avif_encoder_with_internal_color_model(b200avif_encoder *enc, int m) {
  if (!enc || (m != 0 && m != 1)) return fail(B200AVIF_INVALID_ARGUMENT, "unknown color model");
  enc->color_model = m;
  return 0;
}
int b200avif_encoder_with_alpha_color_mode(b200avif_encoder *enc, int m) {
  if (!enc || m < 0 || m > 2) return fail(B200AVIF_INVALID_ARGUMENT, "unknown alpha color mode");
  enc->alpha_color_mode = m;
  enc->premultiplied_alpha = m == B200AVIF_ALPHA_PREMULTIPLIED;   // av1encoder.rs:200-203
  return 0;
}
int b200avif_encoder_with_num_threads(b200avif_encoder *enc, size_t n) {
  if (!enc) return fail(B200AVIF_INVALID_ARGUMENT, "null encoder");
  enc->threads = n;
  return 0;
}
int b200avif_encoder_set_exif(b200avif_encoder *enc, const uint8_t *exif, size_t len) {
  if (!enc || (!exif && len)) return fail(B200AVIF_INVALID_ARGUMENT, "null exif");
  enc->exif.assign(exif, exif + len);
  return 0;
}
int b200avif_encoder_set_device(b200avif_encoder *enc, int dev) {
  if (!enc) return fail(B200AVIF_INVALID_ARGUMENT, "null encoder");
  enc->device = dev;
  return 0;
}

int b200avif_encode_rgba_batch(const b200avif_encoder *enc, const b200avif_input *in, size_t n, b200avif_image *out, int *status) {
  if (!enc || !in || !out) return fail(B200AVIF_INVALID_ARGUMENT, "null argument");
  if (n == 0) return 0;
  std::vector<ImageWork> imgs(n);
  for (size_t i = 0; i < n; i++) {
    ImageWork &im = imgs[i];
    im.pixels = in[i].pixels; im.on_device = in[i].on_device != 0; im.channels = in[i].channels;
    im.w = in[i].width; im.h = in[i].height; im.stride = in[i].stride_px ? in[i].stride_px : in[i].width;
    if (!im.pixels || im.w == 0 || im.h == 0 || (im.channels != 3 && im.channels != 4) || im.stride < im.w) im.status = B200AVIF_INVALID_ARGUMENT;
    if (im.w > 65536 || im.h > 65536) im.status = B200AVIF_UNSUPPORTED;
  }
  return encode_images(*enc, imgs, out, status);
}

int b200avif_encode_rgba(const b200avif_encoder *enc, const uint8_t *rgba, size_t w, size_t h, size_t stride_px, b200avif_image *out) {
  b200avif_input in{rgba, w, h, stride_px, 4, 0};
  int st = 0;
  const int r = b200avif_encode_rgba_batch(enc, &in, 1, out, &st);
  return r ? r : st;
}
int b200avif_encode_rgb(const b200avif_encoder *enc, const uint8_t *rgb, size_t w, size_t h, size_t stride_px, b200avif_image *out) {
  b200avif_input in{rgb, w, h, stride_px, 3, 0};
  int st = 0;
  const int r = b200avif_encode_rgba_batch(enc, &in, 1, out, &st);
  return r ? r : st;
}
int b200avif_encode_raw_planes_8_bit(const b200avif_encoder *enc, size_t w, size_t h, const uint8_t *planes, size_t n_px,
                                     const uint8_t *alpha, size_t n_alpha, int range, int mc, b200avif_image *out) {
  return encode_raw<uint8_t>(enc, w, h, planes, n_px, alpha, n_alpha, range, mc, 8, out);
}
int b200avif_encode_raw_planes_10_bit(const b200avif_encoder *enc, size_t w, size_t h, const uint16_t *planes, size_t n_px,
                                      const uint16_t *alpha, size_t n_alpha, int range, int mc, b200avif_image *out) {
  return encode_raw<uint16_t>(enc, w, h, planes, n_px, alpha, n_alpha, range, mc, 10, out);
}

void b200avif_image_free(b200avif_image *img) {
  if (img && img->avif_file) { free(img->avif_file); img->avif_file = nullptr; img->avif_len = 0; }
}
const char *b200avif_last_error_message(void) { return g_err.c_str(); }
uint8_t b200avif_quality_to_quantizer(float q) { return b200_quality_to_quantizer(q); }

int b200avif_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
size_t b200avif_default_num_threads(void) { return b200_host_threads(); }
void b200_mark(const char *name, void *stream) {
  Context *c = tl_mark_ctx;
  if (!c || !g_profiling.load(std::memory_order_relaxed)) return;
  if (c->mark_used == c->mark_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    c->mark_pool.push_back(e);
  }
  cudaEvent_t e = c->mark_pool[c->mark_used++];
  if (cudaEventRecord(e, (cudaStream_t)stream) != cudaSuccess) return;
  c->marks.push_back(Context::Mark{name, (cudaStream_t)stream, e});
}
size_t b200avif_rate_tables_bytes(void) { return sizeof(B200RateTables); }
int b200avif_build_rate_tables(int base_q_idx, void *host_out) {
  if (!host_out || base_q_idx < 1 || base_q_idx > 255) return fail(B200AVIF_INVALID_ARGUMENT, "bad argument");
  b200av1_build_rate_tables(base_q_idx, (B200RateTables *)host_out);
  return 0;
}
int b200avif_install_rate_tables(int device, int base_q_idx, const void *device_tables, size_t bytes) {
  if (!device_tables || bytes != sizeof(B200RateTables) || base_q_idx < 1 || base_q_idx > 255) return fail(B200AVIF_INVALID_ARGUMENT, "bad argument");
  int cur = 0;
  if (device < 0) { CU(cudaGetDevice(&cur)); device = cur; }
  CU(cudaSetDevice(device));
  B200RateTables *d = nullptr;
  CU(cudaMalloc((void **)&d, sizeof(B200RateTables)));
  CU(cudaMemcpy(d, device_tables, sizeof(B200RateTables), cudaMemcpyDeviceToDevice));
  CU(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_installed_mu);
  auto key = std::make_pair(device, base_q_idx);
  auto it = g_installed_rate.find(key);
  if (it != g_installed_rate.end()) { cudaFree(it->second); it->second = d; } else g_installed_rate[key] = d;
  return 0;
}
void b200avif_set_profiling(int on) { g_profiling.store(on ? 1 : 0); }
size_t b200avif_get_stage_times(b200avif_stage_time *out, size_t cap) {
  std::lock_guard<std::mutex> lk(g_stat_mu);
  size_t n = 0;
  for (const auto &kv : g_stage) {
    if (out && n < cap) {
      memset(&out[n], 0, sizeof(out[n]));
      strncpy(out[n].name, kv.first.c_str(), sizeof(out[n].name) - 1);
      out[n].ms = kv.second.ms; out[n].launches = kv.second.launches;
    }
    n++;
  }
  return n;
}
// Release every pooled context (streams, events, arenas, cached rate tables).  Safe to call at any quiet point; the next
// encode call rebuilds what it needs.  Not registered with atexit(): the CUDA runtime may be torn down first.
void b200avif_shutdown(void) {
  std::vector<std::unique_ptr<Context>> drop;
  { std::lock_guard<std::mutex> lk(g_ctx_mu); drop.swap(g_ctx_free); }
  drop.clear();
}
void b200avif_get_stats(b200avif_stats *out) {
  if (!out) return;
  std::lock_guard<std::mutex> lk(g_stat_mu);
  out->kernel_launches = g_launches; out->h2d_bytes = g_h2d; out->d2h_bytes = g_d2h;
  out->gpu_ms_rdo = g_ms_rdo; out->gpu_ms_total = g_ms_total; out->host_ms_pack = g_ms_pack;
  out->tma_launches = g_tma; out->tma_timeouts = b200avif_device_count() > 0 ? b200_filters_tma_timeouts() : 0;
}

// ---- narrowest seam ------------------------------------------------------------------------------
int b200av1_encode_planes(const b200av1_config *cfg, const uint16_t *const planes[3], const size_t strides[3], uint8_t **obu, size_t *obu_len) {
  if (!cfg || !planes || !strides || !obu || !obu_len) return fail(B200AVIF_INVALID_ARGUMENT, "null argument");
  if (b200avif_device_count() <= 0) return fail(B200AVIF_ENCODING_ERROR, "no CUDA device available (this library has no CPU fallback)");
  FrameJob job;
  if (int e = configure_frame(&job.f, cfg->width, cfg->height, cfg->bit_depth, cfg->monochrome != 0, cfg->quantizer, cfg->speed,
                              cfg->threads, cfg->pixel_range, cfg->has_color_description != 0, cfg->matrix_coefficients)) return e;
  CtxLease lease;
  if (int e = acquire_context(-1, lease.c)) return e;
  Context &c = *lease.c;
  const size_t w = cfg->width, h = cfg->height;
  if (c.dev.reserve(job_dev_bytes(job, true) + job.f.planes * align256(w * h * 2) + (1 << 20)) ||
      c.pin.reserve(job_pin_bytes(job) + job.f.planes * align256(w * h * 2) + (1 << 20))) return B200AVIF_ENCODING_ERROR;
  c.dev.reset(); c.pin.reset();
  CU(cudaEventRecord(c.ev[0], c.stream));
  job_take(job, c.dev, c.pin, true);
  const int cw = job.f.mi_cols * 4, ch = job.f.mi_rows * 4;
  for (int p = 0; p < job.f.planes; p++) {
    uint16_t *stage = c.pin.take<uint16_t>(w * h);
    uint16_t *d = c.dev.take<uint16_t>(w * h);
    ARENA_CHECK(c);
    for (size_t y = 0; y < h; y++) memcpy(stage + y * w, planes[p] + y * strides[p], w * 2);
    CU(cudaMemcpyAsync(d, stage, w * h * 2, cudaMemcpyHostToDevice, c.stream));
    if (int e = b200_pad_plane_launch(d, (int)w, (int)h, (int)w, job.d_src[p], cw, ch, (int)job.stride(), job.f.bit_depth, c.stream)) return fail(B200AVIF_ENCODING_ERROR, cudaGetErrorString((cudaError_t)e));
    g_launches++; g_h2d += w * h * 2;
  }
  std::vector<FrameJob *> jobs{&job};
  if (int e = run_rdo_and_fetch(c, jobs, c.dev, c.pin, false, nullptr)) return e;
  pack_jobs(jobs);
  if (job.status) return fail(B200AVIF_ENCODING_ERROR, "packer failed");
  *obu = (uint8_t *)malloc(job.obu.size());
  memcpy(*obu, job.obu.data(), job.obu.size());
  *obu_len = job.obu.size();
  lease.ok = true;
  return 0;
}

// test seam: compact level streams of one plane -> the dense per-block layout of records.h (what the oracle writes)
static void expand_stream_plane(const FrameJob &j, int p, int16_t *dense) {
  const B200Frame &f = j.f;
  memset(dense, 0, j.n_coef() * 2);
  for (int tr = 0; tr < f.tile_rows; tr++)
    for (int tc = 0; tc < f.tile_cols; tc++) {
      const int r0 = f.tile_row_start_sb[tr] * 16, r1 = std::min(f.tile_row_start_sb[tr + 1] * 16, f.mi_rows);
      const int c0 = f.tile_col_start_sb[tc] * 16, c1 = std::min(f.tile_col_start_sb[tc + 1] * 16, f.mi_cols);
      for (int sr = r0; sr < r1; sr += 16) {
        const int16_t *cur = j.h_stream_base + j.h_stream_off[p][(size_t)(sr >> 4) * f.tile_cols + tc];
        for (int sc = c0; sc < c1; sc += 16)
          for (int z = 0; z < 256; z++) {
            int ur = 0, uc = 0;
            for (int b = 0; b < 4; b++) { uc |= ((z >> (2 * b)) & 1) << b; ur |= ((z >> (2 * b + 1)) & 1) << b; }
            const int mr = sr + ur, mc = sc + uc;
            if (mr >= f.mi_rows || mc >= f.mi_cols) continue;
            const int bs = j.h_mi[(size_t)mr * f.mi_cols + mc].bsize;
            const int n4 = bs == BLOCK_4X4 ? 1 : (bs == BLOCK_8X8 ? 2 : (bs == BLOCK_16X16 ? 4 : 8));
            if ((mr & (n4 - 1)) || (mc & (n4 - 1))) continue;
            const int e = j.h_eob[p][(size_t)mr * f.mi_cols + mc];
            memcpy(dense + b200_coef_offset(&f, mr, mc), cur, (size_t)e * 2);
            cur += e;
          }
      }
    }
}

// test seams exchange uint16 planes with a contiguous stride; the device holds native-width samples with its own stride
static void widen_plane(const uint8_t *dev_copy, size_t bps, size_t dstride, uint16_t *out, size_t w, size_t h) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++)
      out[y * w + x] = bps == 2 ? reinterpret_cast<const uint16_t *>(dev_copy)[y * dstride + x] : dev_copy[y * dstride + x];
}
static void narrow_plane(const uint16_t *in, size_t w, size_t h, size_t bps, size_t dstride, uint8_t *out) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < dstride; x++) {
      const uint16_t v = in[y * w + (x < w ? x : w - 1)];
      if (bps == 2) reinterpret_cast<uint16_t *>(out)[y * dstride + x] = v; else out[y * dstride + x] = (uint8_t)v;
    }
}

// ---- test seam: colour conversion kernel on host pixels -> coded planes (mi-aligned, edge replicated)
int b200av1_alpha_clean_run(const uint8_t *rgba, size_t w, size_t h, size_t stride_px, int mode, uint8_t *out) {
  if (!rgba || !out || w == 0 || h == 0 || (mode != 1 && mode != 2)) return fail(B200AVIF_INVALID_ARGUMENT, "bad argument");
  if (b200avif_device_count() <= 0) return fail(B200AVIF_ENCODING_ERROR, "no CUDA device available (this library has no CPU fallback)");
  CtxLease lease;
  if (int e = acquire_context(-1, lease.c)) return e;
  Context &c = *lease.c;
  const size_t in_bytes = (stride_px * (h - 1) + w) * 4, img_bytes = w * h * 4;
  if (c.dev.reserve(align256(in_bytes) + 2 * align256(img_bytes) + 8192) || c.pin.reserve(4096)) return B200AVIF_ENCODING_ERROR;
  c.dev.reset(); c.pin.reset();
  uint8_t *d_in = c.dev.take<uint8_t>(in_bytes + 16), *d_tmp = c.dev.take<uint8_t>(img_bytes), *d_out = c.dev.take<uint8_t>(img_bytes);
  unsigned long long *d_acc = c.dev.take<unsigned long long>(4);
  CU(cudaMemcpyAsync(d_in, rgba, in_bytes, cudaMemcpyHostToDevice, c.stream));
  if (int e = b200_alpha_clean_launch(d_in, (int)w, (int)h, (int)stride_px, mode, d_tmp, d_out, d_acc, c.stream))
    return fail(B200AVIF_ENCODING_ERROR, cudaGetErrorString((cudaError_t)e));
  g_launches += mode == 2 ? 1 : 3;
  CU(cudaMemcpyAsync(out, d_out, img_bytes, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  lease.ok = true;
  return 0;
}

int b200av1_convert_run(const uint8_t *pixels, int channels, size_t w, size_t h, size_t stride_px, int depth, int color_model,
                        uint16_t *const planes[4] /* y,u,v,alpha (alpha may be NULL) */, int *any_alpha) {
  if (!pixels || !planes || w == 0 || h == 0) return fail(B200AVIF_INVALID_ARGUMENT, "null argument");
  if (b200avif_device_count() <= 0) return fail(B200AVIF_ENCODING_ERROR, "no CUDA device available (this library has no CPU fallback)");
  CtxLease lease;
  if (int e = acquire_context(-1, lease.c)) return e;
  Context &c = *lease.c;
  const int cw = 8 * (int)((w + 7) >> 3), ch = 8 * (int)((h + 7) >> 3);
  const size_t npx = (size_t)cw * ch, in_bytes = (stride_px * (h - 1) + w) * channels;
  if (c.dev.reserve(4 * align256(npx * 2) + align256(in_bytes) + 8192) || c.pin.reserve(4096)) return B200AVIF_ENCODING_ERROR;
  c.dev.reset(); c.pin.reset();
  uint8_t *d_in = c.dev.take<uint8_t>(in_bytes + 16);
  const size_t bps = depth > 8 ? 2 : 1;
  uint8_t *d_p[4];
  for (int p = 0; p < 4; p++) d_p[p] = c.dev.take<uint8_t>(npx * bps);
  int *d_flag = c.dev.take<int>(1);
  CU(cudaMemcpyAsync(d_in, pixels, in_bytes, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemsetAsync(d_flag, 0, sizeof(int), c.stream));
  if (int e = b200_convert_launch(d_in, channels, (int)w, (int)h, (int)stride_px, cw, ch, cw, depth, color_model, d_p[0], d_p[1], d_p[2],
                                  channels == 4 ? d_p[3] : nullptr, channels == 4 ? d_flag : nullptr, c.stream))
    return fail(B200AVIF_ENCODING_ERROR, cudaGetErrorString((cudaError_t)e));
  g_launches++;
  std::vector<uint8_t> host[4];
  for (int p = 0; p < 4; p++)
    if (planes[p] && (p < 3 || channels == 4)) { host[p].resize(npx * bps); CU(cudaMemcpyAsync(host[p].data(), d_p[p], npx * bps, cudaMemcpyDeviceToHost, c.stream)); }
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  for (int p = 0; p < 4; p++) if (!host[p].empty()) widen_plane(host[p].data(), bps, (size_t)cw, planes[p], (size_t)cw, (size_t)ch);
  if (any_alpha) *any_alpha = flag;
  lease.ok = true;
  return 0;
}

// ---- test seam: run the device hot path on caller-provided coded planes and frame description -----
int b200av1_rdo_run(const B200Frame *f, const uint16_t *const src[3], B200Mi *mi, uint16_t *const eob[3], int16_t *const coef[3],
                    uint16_t *const rec[3], uint16_t *const deblocked[3], uint16_t *const fin[3], int8_t *cdef_idx,
                    B200BlockDec *const dec[B200_LEVELS], uint32_t *const satd[B200_LEVELS], B200LrUnit *const lr[3], uint16_t *const lr_out[3],
                    int32_t *lf_level_out /* may be NULL: the four deblocking levels the frame ends up with */) {
  if (!f || !src || !mi) return fail(B200AVIF_INVALID_ARGUMENT, "null argument");
  if (b200avif_device_count() <= 0) return fail(B200AVIF_ENCODING_ERROR, "no CUDA device available (this library has no CPU fallback)");
  FrameJob job;
  job.f = *f;
  CtxLease lease;
  if (int e = acquire_context(-1, lease.c)) return e;
  Context &c = *lease.c;
  if (c.dev.reserve(job_dev_bytes(job, true) + 9 * align256(job.plane_bytes()) + (1 << 20)) || c.pin.reserve(job_pin_bytes(job) + 3 * align256(job.plane_bytes()) + (1 << 20))) return B200AVIF_ENCODING_ERROR;
  c.dev.reset(); c.pin.reset();
  CU(cudaEventRecord(c.ev[0], c.stream));
  job_take(job, c.dev, c.pin, true);
  for (int p = 0; p < job.f.planes; p++) {
    if (!job.d_fin[p]) job.d_fin[p] = c.dev.take<uint8_t>(job.plane_bytes());
    job.d_deb_full[p] = c.dev.take<uint8_t>(job.plane_bytes());
    if (job.uses_lr() && lr_out && lr_out[p]) job.d_lr_out[p] = c.dev.take<uint8_t>(job.plane_bytes());
    uint8_t *stage = c.pin.take<uint8_t>(job.plane_bytes());
    ARENA_CHECK(c);
    narrow_plane(src[p], (size_t)job.f.mi_cols * 4, (size_t)job.f.mi_rows * 4, job.bps(), job.stride(), stage);
    CU(cudaMemcpyAsync(job.d_src[p], stage, job.plane_bytes(), cudaMemcpyHostToDevice, c.stream));
    // test seam only: slots of blocks outside the frame are never written (and never read by the packer);
    // zero them so that whole-buffer comparisons against the oracle are meaningful
    CU(cudaMemsetAsync(job.d_coef[p], 0, job.n_coef() * 2, c.stream));
    CU(cudaMemsetAsync(job.d_eob[p], 0, job.n_mi() * 2, c.stream));
  }
  for (int lv = 0; lv < B200_LEVELS; lv++) {   // levels outside the partition range are not analysed: keep them comparable
    CU(cudaMemsetAsync(job.d_dec[lv], 0, job.n_blk(lv) * sizeof(B200BlockDec), c.stream));
    CU(cudaMemsetAsync(job.d_satd[lv], 0, job.n_blk(lv) * B200_SATD_WORDS * 4, c.stream));
  }
  std::vector<FrameJob *> jobs{&job};
  std::vector<std::vector<uint8_t>> deb, pre;
  if (int e = run_rdo_and_fetch(c, jobs, c.dev, c.pin, true, &deb, &pre)) return e;
  memcpy(mi, job.h_mi, job.n_mi() * sizeof(B200Mi));
  if (lf_level_out) for (int k = 0; k < 4; k++) lf_level_out[k] = job.f.lf_search ? job.h_lf_out[k] : job.f.lf_level[k];
  if (cdef_idx) memcpy(cdef_idx, job.h_cdef, job.n_sb());
  for (int lv = 0; lv < B200_LEVELS; lv++) {
    if (dec && dec[lv]) CU(cudaMemcpy(dec[lv], job.d_dec[lv], job.n_blk(lv) * sizeof(B200BlockDec), cudaMemcpyDeviceToHost));
    if (satd && satd[lv]) CU(cudaMemcpy(satd[lv], job.d_satd[lv], job.n_blk(lv) * B200_SATD_WORDS * 4, cudaMemcpyDeviceToHost));
  }
  for (int p = 0; p < job.f.planes; p++) {
    if (eob && eob[p]) memcpy(eob[p], job.h_eob[p], job.n_mi() * 2);
    if (coef && coef[p]) expand_stream_plane(job, p, coef[p]);
    const size_t cw = (size_t)job.f.mi_cols * 4, chh = (size_t)job.f.mi_rows * 4;
    std::vector<uint8_t> tmp(job.plane_bytes());
    if (rec && rec[p]) widen_plane(pre[p].data(), job.bps(), job.stride(), rec[p], cw, chh);
    if (deblocked && deblocked[p]) widen_plane(deb[p].data(), job.bps(), job.stride(), deblocked[p], cw, chh);
    if (fin && fin[p]) { CU(cudaMemcpy(tmp.data(), job.d_fin[p], job.plane_bytes(), cudaMemcpyDeviceToHost)); widen_plane(tmp.data(), job.bps(), job.stride(), fin[p], cw, chh); }
    if (job.uses_lr() && lr && lr[p]) memcpy(lr[p], job.h_lr[p], job.n_lr() * sizeof(B200LrUnit));
    if (job.d_lr_out[p]) { CU(cudaMemcpy(tmp.data(), job.d_lr_out[p], job.plane_bytes(), cudaMemcpyDeviceToHost)); widen_plane(tmp.data(), job.bps(), job.stride(), lr_out[p], cw, chh); }
  }
  lease.ok = true;
  return 0;
}

}  // extern "C"
