// RGBA8 / RGB8 (interleaved, strided) -> planar 8/10-bit Y,Cb,Cr (or G,B,R) + alpha plane,
// written at the NATIVE sample width (uint8_t planes for 8-bit, uint16_t for 10-bit) into 8-pixel-aligned,
// edge-replicated coded planes (row stride `out_stride` samples), plus an
// "any pixel with a != 255" flag.  One streaming pass, HBM-bound.
//
// Bit-exact restatement of the reference's per-pixel arithmetic:
//   rgb_to_ycbcr            /root/reference/ravif/src/av1encoder.rs:504-512 (f32, explicit FMAs, round())
//   rgb_to_{8,10}_bit_ycbcr :514-524 (`as u8` saturates, `as u16` does not clamp to 1023)
//   rgb_to_{8,10}_bit_gbr / to_ten  :485-498
//   use_alpha = any(a != 255)       :246
// and of init_frame_3/_1 (:710-746), which scatter the samples row-major into the frame planes.
// The FMA placement matters for 569 of 2^24 inputs at 10 bit, so every float
// operation is an explicit round-to-nearest intrinsic (never contracted by the compiler).
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

struct ConvParams {
  const uint8_t *in;      // device, interleaved
  int channels;           // 3 | 4
  int width, height;      // visible
  int stride_px;          // input row stride in pixels
  int cw, ch;             // coded (padded) plane size
  int out_stride;         // samples per row of the output planes (>= cw, multiple of 4)
  int depth;              // 8 | 10
  int color_model;        // 0 YCbCr(BT601) 1 RGB (planes G,B,R)
  void *p0, *p1, *p2, *alpha;   // alpha may be null; uint8_t (depth 8) or uint16_t (depth 10) samples
  int *any_alpha;         // device flag (may be null)
};

__device__ __forceinline__ uint16_t to_ten(uint8_t x) { return (uint16_t)(((uint16_t)x << 2) | ((uint16_t)x >> 6)); }

__device__ __forceinline__ void convert_px(uint8_t r, uint8_t g, uint8_t b, int depth, int color_model, uint16_t o[3]) {
  if (color_model == 1) {
    if (depth == 8) { o[0] = g; o[1] = b; o[2] = r; }
    else { o[0] = to_ten(g); o[1] = to_ten(b); o[2] = to_ten(r); }
    return;
  }
  const float KR = 0.2990f, KG = 0.5870f, KB = 0.1140f;
  const float max_value = (float)((1 << depth) - 1);
  const float scale = __fdiv_rn(max_value, 255.f);
  const float shift = roundf(__fmul_rn(max_value, 0.5f));
  const float fr = (float)r, fg = (float)g, fb = (float)b;
  const float y = __fmaf_rn(__fmul_rn(scale, KB), fb, __fmaf_rn(__fmul_rn(scale, KR), fr, __fmul_rn(__fmul_rn(scale, KG), fg)));
  const float cb = __fmaf_rn(__fmaf_rn(fb, scale, -y), __fdiv_rn(0.5f, __fsub_rn(1.f, KB)), shift);
  const float cr = __fmaf_rn(__fmaf_rn(fr, scale, -y), __fdiv_rn(0.5f, __fsub_rn(1.f, KR)), shift);
  const float v[3] = {roundf(y), roundf(cb), roundf(cr)};
  const float hi = depth == 8 ? 255.f : 65535.f;
#pragma unroll
  for (int k = 0; k < 3; k++) o[k] = (uint16_t)(v[k] <= 0.f ? 0.f : (v[k] >= hi ? hi : v[k]));
}

template <class PIX> struct Vec4;
template <> struct Vec4<uint8_t> { typedef uchar4 type; };
template <> struct Vec4<uint16_t> { typedef ushort4 type; };

// One thread per FOUR horizontally adjacent coded pixels (the coded width is a multiple of 8): one 16-byte load of
// RGBA8 where the row is 16-byte aligned, one 4- or 8-byte store per plane.  Threads beyond the visible area replicate
// the last column / row.
template <class PIX>
__global__ void __launch_bounds__(256) convert_kernel(ConvParams p) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  if (x4 >= p.cw) return;
  const int sy = min(y, p.height - 1);
  uint8_t r[4], g[4], b[4], a[4];
  const uint8_t *row = p.in + (size_t)sy * p.stride_px * p.channels;
  if (p.channels == 4 && x4 + 3 < p.width && ((reinterpret_cast<uintptr_t>(row) + (size_t)x4 * 4) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *>(row + (size_t)x4 * 4);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) { r[k] = w[k] & 255; g[k] = (w[k] >> 8) & 255; b[k] = (w[k] >> 16) & 255; a[k] = w[k] >> 24; }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint8_t *px = row + (size_t)min(x4 + k, p.width - 1) * p.channels;
      if (p.channels == 4) {
        const uchar4 v = *reinterpret_cast<const uchar4 *>(px);
        r[k] = v.x; g[k] = v.y; b[k] = v.z; a[k] = v.w;
      } else {
        r[k] = px[0]; g[k] = px[1]; b[k] = px[2]; a[k] = 255;
      }
    }
  }
  uint16_t o[4][3];
#pragma unroll
  for (int k = 0; k < 4; k++) convert_px(r[k], g[k], b[k], p.depth, p.color_model, o[k]);
  typedef typename Vec4<PIX>::type V4;
  const size_t idx = ((size_t)y * p.out_stride + x4) >> 2;
  V4 v;
  v.x = (PIX)o[0][0]; v.y = (PIX)o[1][0]; v.z = (PIX)o[2][0]; v.w = (PIX)o[3][0]; reinterpret_cast<V4 *>(p.p0)[idx] = v;
  v.x = (PIX)o[0][1]; v.y = (PIX)o[1][1]; v.z = (PIX)o[2][1]; v.w = (PIX)o[3][1]; reinterpret_cast<V4 *>(p.p1)[idx] = v;
  v.x = (PIX)o[0][2]; v.y = (PIX)o[1][2]; v.z = (PIX)o[2][2]; v.w = (PIX)o[3][2]; reinterpret_cast<V4 *>(p.p2)[idx] = v;
  if (p.alpha) {
    if (sizeof(PIX) == 1) { v.x = (PIX)a[0]; v.y = (PIX)a[1]; v.z = (PIX)a[2]; v.w = (PIX)a[3]; }
    else { v.x = (PIX)to_ten(a[0]); v.y = (PIX)to_ten(a[1]); v.z = (PIX)to_ten(a[2]); v.w = (PIX)to_ten(a[3]); }
    reinterpret_cast<V4 *>(p.alpha)[idx] = v;
  }
  if (p.any_alpha && y < p.height) {
    bool t = false;
#pragma unroll
    for (int k = 0; k < 4; k++) t |= (x4 + k < p.width) && a[k] != 255;
    const unsigned m = __ballot_sync(__activemask(), t);
    if (m && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicOr(p.any_alpha, 1);
  }
}

// raw planar input (encode_raw_planes_*): copy + edge-replicate into the coded plane (narrowing to the native width)
template <class PIX>
__global__ void __launch_bounds__(256) pad_plane_kernel(const uint16_t *in, int width, int height, int in_stride, PIX *out, int cw, int out_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= cw) return;
  out[(size_t)y * out_stride + x] = (PIX)in[(size_t)min(y, height - 1) * in_stride + min(x, width - 1)];
}

}  // namespace

extern "C" int b200_convert_launch(const uint8_t *d_in, int channels, int width, int height, int stride_px, int cw, int ch, int out_stride,
                                   int depth, int color_model, void *p0, void *p1, void *p2, void *alpha,
                                   int *any_alpha, void *stream) {
  ConvParams p{d_in, channels, width, height, stride_px, cw, ch, out_stride, depth, color_model, p0, p1, p2, alpha, any_alpha};
  dim3 grid((cw / 4 + 255) / 256, ch);
  if (depth > 8) convert_kernel<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else convert_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int b200_pad_plane_launch(const uint16_t *d_in, int width, int height, int in_stride, void *d_out, int cw, int ch, int out_stride,
                                     int depth, void *stream) {
  dim3 grid((cw + 255) / 256, ch);
  if (depth > 8) pad_plane_kernel<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, width, height, in_stride, (uint16_t *)d_out, cw, out_stride);
  else pad_plane_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, width, height, in_stride, (uint8_t *)d_out, cw, out_stride);
  return (int)cudaGetLastError();
}
