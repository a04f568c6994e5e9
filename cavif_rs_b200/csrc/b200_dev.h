// Internal device-side entry points (kernel launchers) and the test seam of the hot path.
#ifndef B200_DEV_H_
#define B200_DEV_H_
#include <stddef.h>
#include <stdint.h>
#include "records.h"
#ifdef __cplusplus
extern "C" {
#endif
// output planes: native sample width (uint8_t for depth 8, uint16_t for depth 10), out_stride samples per row
int b200_convert_launch(const uint8_t *d_in, int channels, int width, int height, int stride_px, int cw, int ch, int out_stride, int depth,
                        int color_model, void *p0, void *p1, void *p2, void *alpha, int *any_alpha,
                        void *stream);
// alpha_clean.cu: dirtyalpha.rs / premultiplied branch on the device (mode 1 = UnassociatedClean, 2 = Premultiplied)
int b200_alpha_clean_launch(const uint8_t *d_in, int w, int h, int stride_px, int mode, uint8_t *d_tmp, uint8_t *d_out,
                            unsigned long long *d_acc, void *stream);
int b200_pad_plane_launch(const uint16_t *d_in, int width, int height, int in_stride, void *d_out, int cw, int ch, int out_stride,
                          int depth, void *stream);
// Test seam: colour-conversion kernel.  planes[p]: (8*ceil(h/8)) x (8*ceil(w/8)) uint16 each.
// test seam: runs the device alpha pre-clean on host pixels, returns the cleaned w*h RGBA8 image
int b200av1_alpha_clean_run(const uint8_t *rgba, size_t w, size_t h, size_t stride_px, int mode, uint8_t *out);
int b200av1_convert_run(const uint8_t *pixels, int channels, size_t w, size_t h, size_t stride_px, int depth, int color_model,
                        uint16_t *const planes[4], int *any_alpha);
// Test seam: run the device hot path (RDO kernel, then deblock + CDEF search) on host-provided CODED planes
// (mi_rows*4 x mi_cols*4 uint16, stride mi_cols*4).  Returns the block records, the reconstruction before the
// in-loop filters (rec), after deblocking (deblocked), after CDEF (fin), the per-superblock CDEF index and the
// open-loop analysis outputs per level (decisions, SATD tables).
int b200av1_rdo_run(const B200Frame *f, const uint16_t *const src[3], B200Mi *mi, uint16_t *const eob[3],
                    int16_t *const coef[3], uint16_t *const rec[3], uint16_t *const deblocked[3], uint16_t *const fin[3],
                    int8_t *cdef_idx, B200BlockDec *const dec[B200_LEVELS], uint32_t *const satd[B200_LEVELS], B200LrUnit *const lr[3], uint16_t *const lr_out[3],
                    int32_t *lf_level_out);
#ifdef __cplusplus
}
#endif
#endif
