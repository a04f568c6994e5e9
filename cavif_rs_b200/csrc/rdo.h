// Device job descriptors of the RDO kernel (rdo.cu).
#ifndef B200_RDO_H_
#define B200_RDO_H_
#include <stdint.h>
#include "records.h"
#ifdef __cplusplus
extern "C" {
#endif

// One AV1 frame to encode (the colour planes of an image, or its alpha plane): all pointers are
// device pointers.  Planes hold samples at their NATIVE width - uint8_t for 8-bit video (frame.bit_depth == 8),
// uint16_t for 10-bit - with `stride` samples per row (>= mi_cols*4) and mi_rows*4 rows.
typedef struct B200RdoJob {
  B200Frame frame;
  const B200RateTables *rate;
  int32_t stride;          // samples per row of src / rec / fin / lr_out
  int32_t reserved_;
  const void *src[3];
  void *rec[3];
  B200Mi *mi;
  uint16_t *eob[3];
  int16_t *coef[3];
  // compact level streams (recon.cu): length of the stream of every (frame superblock row, tile column) region per plane,
  // [sb_rows * tile_cols]; the stream itself starts at the region's first slot of coef[plane]
  uint32_t *stream_len[3];
  B200BlockDec *dec[B200_LEVELS];   // analysis decisions per level (0: 4x4 .. 3: 32x32), b200_level_rows*cols entries
  uint32_t *satd[B200_LEVELS];      // luma SATD tables per level: [block][14 words]
  // survivors of the top-down pruning per level: block indices appended by the SATD kernel (any order), count in list_n[lv];
  // the lock-step analysis of the pruned levels walks these lists
  uint32_t *list[B200_LEVELS];
  uint32_t *list_n;                 // [B200_LEVELS], zeroed by the top-level SATD kernel
  // in-loop filter stage (filters.cu): rec is read only (deblocking happens on a staged copy); cdef_idx receives the per-superblock
  // strength index (-1: nothing coded in the superblock); fin (optional, may be NULL) receives the CDEF output.
  int8_t *cdef_idx;       // sb_rows*sb_cols
  void *fin[3];
  // deblocking level search (frame.lf_search): squared error of every (plane, candidate) accumulated by the search kernel,
  // [3][B200_LF_CANDIDATES], zeroed by the caller; lf_out receives the four chosen levels (the kernel also writes frame.lf_level)
  unsigned long long *lf_sse;
  int32_t *lf_out;
  void *dbrows[3];        // deblocked rows 54..57 of every superblock row (4 * sb_rows rows of `stride` samples): loop restoration input
  void *deb_full[3];      // optional (tests): the whole deblocked planes
  // loop restoration (frame.lr_type[p] != 0; needs fin): per-unit decisions, and (optional, tests) the restored planes
  B200LrUnit *lr[3];      // b200_lr_units(height) * b200_lr_units(width) per plane
  void *lr_out[3];
} B200RdoJob;

typedef struct B200TileRef {
  int32_t job;
  int16_t tile_row, tile_col;
} B200TileRef;

// Launch one CTA per tile on `stream` (cudaStream_t).  Returns a cudaError_t value.
// All RDO stages for all jobs on `stream`: SATD + analysis per level, partition decision, reconstruction chain
// (one warp per tile).  max_blocks4 = largest mi_rows*mi_cols over the jobs; high_bitdepth: all jobs are 10-bit.
int b200_rdo_launch(const B200RdoJob *d_jobs, int n_jobs, const B200TileRef *d_tiles, int n_tiles, int max_blocks4,
                    int max_regions /* largest root-region count (ceil(W/rn) * ceil(H/rn), rn = 16 or 32) over the jobs */, int high_bitdepth, int top_level /* 2: 16x16, 3: 32x32 */, int recon_rows /* superblock-row slots per tile: 1, 2, 4 */, void *stream, void *stream_recon, void *analysis_done,
                    int *launched /* out, may be NULL: kernels launched by this call */);
// Per-kernel timing hook (api.cpp): the launchers call it after every kernel; when profiling is on it records a CUDA event
// on that stream, and the time between consecutive marks of a stream is accumulated under `name`
// (b200avif_get_stage_times).  name == NULL opens an interval without closing one.
void b200_mark(const char *name, void *stream);
// recon.cu: the reconstruction chain (one CTA per tile) and the gather of the compact level streams
int b200_recon_launch(const B200RdoJob *d_jobs, const B200TileRef *d_tiles, int n_tiles, int high_bitdepth, int top_level, int rows, void *stream);
size_t b200_recon_smem_bytes(int high_bitdepth, int top_level, int rows);
typedef struct B200StreamRegionH { const int16_t *src; const uint32_t *len; } B200StreamRegionH;   // device pointers
int b200_stream_gather_launch(const void *d_regions, int n_regions, uint32_t *d_offsets /* n + 1 */, int16_t *out, uint32_t capacity, void *stream);
// Fused deblocking + CDEF search (1 launch) + loop-restoration search (1 launch when with_lr) for all jobs: grid.y = job.
// d_tmaps: CUtensorMap[n_jobs][3] of the rec planes (box from b200_filters_tmap_box; coordinates in samples), or NULL
// with_lf_search: some job has frame.lf_search set (two more launches: level search + pick, in front of the fused kernel)
int b200_filters_launch(B200RdoJob *d_jobs, int n_jobs, int max_sb, int with_lr, int with_lf_search, int high_bitdepth, const void *d_tmaps,
                        int tma_x_shift /* 1: 8-bit planes described as 16-bit elements */, void *stream, int *launched);
void b200_filters_tmap_box(int high_bitdepth, int *box_w, int *box_h);
unsigned b200_filters_tma_timeouts(void);   // TMA tiles that never landed on the current device (0 in a healthy run)

#ifdef __cplusplus
}
#endif
#endif
