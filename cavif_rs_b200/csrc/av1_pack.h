// Host-side packer interface (records -> AV1 OBUs).  See av1_pack.cpp.
#ifndef B200_AV1_PACK_H_
#define B200_AV1_PACK_H_
#include <stddef.h>
#include <stdint.h>
#include "records.h"

#ifdef __cplusplus
extern "C" {
#endif

// Host views of what the hot path produced for one frame (colour: 3 planes, alpha: 1 plane).
typedef struct B200Records {
  const B200Mi *mi;          // [mi_rows * mi_cols]
  const uint16_t *eob[3];    // per plane [mi_rows * mi_cols], valid at the top-left unit of each transform block
  const int16_t *coef[3];    // per plane [sb_rows * sb_cols * 4096], quantised levels in scan order (b200_coef_offset)
  // Compact level streams (device path, recon.cu); when stream_base != NULL `coef` is not read.  The eob levels of every
  // transform block (scan order) of region (plane p, frame superblock row r, tile column c), appended in coding order,
  // start at stream_base + stream_off[p][r * tile_cols + c].
  const int16_t *stream_base;
  const uint32_t *stream_off[3];
  const int8_t *cdef_idx;    // [sb_rows * sb_cols] or NULL
  const B200LrUnit *lr[3];   // per plane [b200_lr_units(height) * b200_lr_units(width)]; read only where f->lr_type[plane] != 0
} B200Records;

// Fill geometry/derived fields of B200Frame (mi/sb sizes, uniform tile layout, quantiser steps, lambda).
// tiles_target: desired tile count (ravif: min(threads, w*h/min_tile_size^2), av1encoder.rs:665-668); 0/1 = single tile.
int b200av1_frame_setup(B200Frame *f, int width, int height, int bit_depth, int planes, int quantizer,
                        int tiles_target);

// Static RDO rate model for the quantiser class of base_q_idx.
void b200av1_build_rate_tables(int base_q_idx, B200RateTables *t);

// seq_level_idx_0 written into the sequence header (needed by the av1C box).
// Host threads b200av1_pack_frame may use for the tiles of one frame (default 1; batches already pack one frame per thread).
void b200av1_pack_set_tile_threads(int n);
int b200av1_seq_level_idx(const B200Frame *f);

// Pack one frame.  *out is malloc()ed; caller frees with free().  Returns 0 on success.
int b200av1_pack_frame(const B200Frame *f, const B200Records *rec, uint8_t **out, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
