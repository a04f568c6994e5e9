// RGBA8 / RGB8 (interleaved, strided) -> planar 8/10-bit Y,Cb,Cr (or G,B,R) + alpha plane,
// written as uint16 samples into 8-pixel-aligned, edge-replicated coded planes, plus an
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
  int depth;              // 8 | 10
  int color_model;        // 0 YCbCr(BT601) 1 RGB (planes G,B,R)
  uint16_t *p0, *p1, *p2, *alpha;   // alpha may be null
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

// one thread per coded pixel; threads beyond the visible area replicate the last column/row
__global__ void __launch_bounds__(256) convert_kernel(ConvParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.cw) return;
  const int sx = min(x, p.width - 1), sy = min(y, p.height - 1);
  const uint8_t *px = p.in + ((size_t)sy * p.stride_px + sx) * p.channels;
  uint8_t r, g, b, a = 255;
  if (p.channels == 4) {
    const uchar4 v = *reinterpret_cast<const uchar4 *>(px);
    r = v.x; g = v.y; b = v.z; a = v.w;
  } else {
    r = px[0]; g = px[1]; b = px[2];
  }
  uint16_t o[3];
  convert_px(r, g, b, p.depth, p.color_model, o);
  const size_t idx = (size_t)y * p.cw + x;
  p.p0[idx] = o[0]; p.p1[idx] = o[1]; p.p2[idx] = o[2];
  if (p.alpha) p.alpha[idx] = p.depth == 8 ? (uint16_t)a : to_ten(a);
  if (p.any_alpha && x < p.width && y < p.height) {
    const unsigned m = __ballot_sync(__activemask(), a != 255);
    if (m && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicOr(p.any_alpha, 1);
  }
}

// raw planar input (encode_raw_planes_*): copy + edge-replicate into the coded plane
__global__ void __launch_bounds__(256) pad_plane_kernel(const uint16_t *in, int width, int height, int in_stride, uint16_t *out, int cw) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= cw) return;
  out[(size_t)y * cw + x] = in[(size_t)min(y, height - 1) * in_stride + min(x, width - 1)];
}

}  // namespace

extern "C" int b200_convert_launch(const uint8_t *d_in, int channels, int width, int height, int stride_px, int cw, int ch,
                                   int depth, int color_model, uint16_t *p0, uint16_t *p1, uint16_t *p2, uint16_t *alpha,
                                   int *any_alpha, void *stream) {
  ConvParams p{d_in, channels, width, height, stride_px, cw, ch, depth, color_model, p0, p1, p2, alpha, any_alpha};
  dim3 grid((cw + 255) / 256, ch);
  convert_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int b200_pad_plane_launch(const uint16_t *d_in, int width, int height, int in_stride, uint16_t *d_out, int cw, int ch,
                                     void *stream) {
  dim3 grid((cw + 255) / 256, ch);
  pad_plane_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, width, height, in_stride, d_out, cw);
  return (int)cudaGetLastError();
}
