// Alpha pre-clean of RGBA8 input on the device: the colour under transparent pixels is replaced by
// a blur of the nearby visible colour so that it costs no bits.  Bit-exact restatement of
//   blurred_dirty_alpha        /root/reference/ravif/src/dirtyalpha.rs:17-43  (weighted average of "edge" pixels; None when
//                                                                            nothing semi-transparent touches a=0)
//   bleed_opaque_color         :45-76   (3x3 weighted average, weight 256 - a, clamp-to-edge neighbourhood)
//   blur_transparent_pixels    :79-100  (3x3 box blur of the non-opaque pixels)
//   premultiplied_minmax/clamp :103-124 (keeps semi-transparent colours within the range that premultiplication rounds to)
// and of the premultiplied branch of convert_alpha_8bit (av1encoder.rs:277-299).
// Three streaming kernels (one reduction + two 3x3 stencils), all HBM-bound; pixels are uchar4 loads/stores.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

struct CleanParams {
  const uint8_t *in;   // RGBA8, stride_px pixels per row
  int w, h, stride_px;
  uint8_t *tmp;        // w*h RGBA8
  uint8_t *out;        // w*h RGBA8
  unsigned long long *acc;   // [4]: sum r, g, b of weighted edge pixels, sum of weights
};

__device__ __forceinline__ uchar4 px_at(const uint8_t *p, int w, int h, int stride, int x, int y) {
  x = min(max(x, 0), w - 1);
  y = min(max(y, 0), h - 1);
  return *reinterpret_cast<const uchar4 *>(p + ((size_t)y * stride + x) * 4);
}
__device__ __forceinline__ uint32_t weighed(uchar4 p, uint32_t rgb[3]) {   // dirtyalpha.rs:5-14
  if (p.w == 0) { rgb[0] = rgb[1] = rgb[2] = 0; return 0; }
  const uint32_t w = 256u - p.w;
  rgb[0] = p.x * w; rgb[1] = p.y * w; rgb[2] = p.z * w;
  return w;
}
__device__ __forceinline__ uint8_t clamp_minmax(uint8_t v, uint8_t px, uint8_t alpha) {   // dirtyalpha.rs:103-124
  const uint32_t a = alpha;
  const uint32_t rounded = ((uint32_t)px * a / 255u * 255u) & 0xffffu;
  const uint8_t low = (uint8_t)((rounded + 16u) / a), hi = (uint8_t)((rounded + 239u) / a);
  const uint8_t lo2 = low < px ? low : px, hi2 = hi > px ? hi : px;
  const uint8_t t = v > lo2 ? v : lo2;
  return t < hi2 ? t : hi2;
}

// pass 1: weighted sum of the semi-transparent pixels that touch a fully transparent one
__global__ void __launch_bounds__(256) dirty_sum_kernel(CleanParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  unsigned long long s[4] = {0, 0, 0, 0};
  if (x < p.w) {
    const uchar4 c = px_at(p.in, p.w, p.h, p.stride_px, x, y);
    if (c.w != 255 && c.w != 0) {
      bool near0 = false;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) near0 |= px_at(p.in, p.w, p.h, p.stride_px, x + dx, y + dy).w == 0;
      if (near0) {
        uint32_t rgb[3];
        s[3] = weighed(c, rgb);
        s[0] = rgb[0]; s[1] = rgb[1]; s[2] = rgb[2];
      }
    }
  }
  // rows of opaque pixels (the common case) leave without touching the accumulators
  if (!__syncthreads_or(s[3] != 0)) return;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    for (int m = 16; m > 0; m >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], m);
    if ((threadIdx.x & 31) == 0 && s[k]) atomicAdd(p.acc + k, s[k]);
  }
}

// pass 2: bleed_opaque_color (in -> tmp)
__global__ void __launch_bounds__(256) bleed_kernel(CleanParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.w) return;
  const unsigned long long weights = p.acc[3];
  if (weights == 0) return;            // blurred_dirty_alpha() returned None: pass 3 copies the input through
  const uchar4 c = px_at(p.in, p.w, p.h, p.stride_px, x, y);
  uchar4 o = c;
  if (c.w != 255) {
    uint32_t ws = 0, s[3] = {0, 0, 0};
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        uint32_t rgb[3];
        ws += weighed(px_at(p.in, p.w, p.h, p.stride_px, x + dx, y + dy), rgb);
        s[0] += rgb[0]; s[1] += rgb[1]; s[2] += rgb[2];
      }
    if (ws == 0) {
      o = make_uchar4((uint8_t)(p.acc[0] / weights), (uint8_t)(p.acc[1] / weights), (uint8_t)(p.acc[2] / weights), 0);
    } else {
      const uint8_t a0 = (uint8_t)(s[0] / ws), a1 = (uint8_t)(s[1] / ws), a2 = (uint8_t)(s[2] / ws);
      if (c.w == 0) o = make_uchar4(a0, a1, a2, 0);
      else o = make_uchar4(clamp_minmax(a0, c.x, c.w), clamp_minmax(a1, c.y, c.w), clamp_minmax(a2, c.z, c.w), c.w);
    }
  }
  *reinterpret_cast<uchar4 *>(p.tmp + ((size_t)y * p.w + x) * 4) = o;
}

// pass 3: blur_transparent_pixels (tmp -> out), or the untouched input when there was nothing to clean
__global__ void __launch_bounds__(256) blur_kernel(CleanParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.w) return;
  uchar4 o;
  if (p.acc[3] == 0) {
    o = px_at(p.in, p.w, p.h, p.stride_px, x, y);
  } else {
    const uchar4 c = px_at(p.tmp, p.w, p.h, p.w, x, y);
    o = c;
    if (c.w != 255) {
      uint32_t s[3] = {0, 0, 0};
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          const uchar4 n = px_at(p.tmp, p.w, p.h, p.w, x + dx, y + dy);
          s[0] += n.x; s[1] += n.y; s[2] += n.z;
        }
      const uint8_t a0 = (uint8_t)(s[0] / 9u), a1 = (uint8_t)(s[1] / 9u), a2 = (uint8_t)(s[2] / 9u);
      if (c.w == 0) o = make_uchar4(a0, a1, a2, 0);
      else o = make_uchar4(clamp_minmax(a0, c.x, c.w), clamp_minmax(a1, c.y, c.w), clamp_minmax(a2, c.z, c.w), c.w);
    }
  }
  *reinterpret_cast<uchar4 *>(p.out + ((size_t)y * p.w + x) * 4) = o;
}

// AlphaColorMode::Premultiplied (av1encoder.rs:277-299): c*255/a in u16, truncated to the low byte; a in {0, 255} -> 0
__global__ void __launch_bounds__(256) premultiplied_kernel(CleanParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.w) return;
  const uchar4 c = px_at(p.in, p.w, p.h, p.stride_px, x, y);
  uchar4 o = make_uchar4(0, 0, 0, 0);
  if (c.w != 0 && c.w != 255) {
    const uint32_t a = c.w;
    o = make_uchar4((uint8_t)(((uint32_t)c.x * 255u / a) & 0xffffu), (uint8_t)(((uint32_t)c.y * 255u / a) & 0xffffu),
                    (uint8_t)(((uint32_t)c.z * 255u / a) & 0xffffu), c.w);
  }
  *reinterpret_cast<uchar4 *>(p.out + ((size_t)y * p.w + x) * 4) = o;
}

}  // namespace

// mode: 1 = UnassociatedClean, 2 = Premultiplied (UnassociatedDirty never calls this).  d_acc: 4 x u64 scratch,
// d_tmp / d_out: w*h*4 bytes each (d_tmp unused for mode 2).  Returns 0 or a cudaError_t; 3 (4 for mode 1) launches.
extern "C" int b200_alpha_clean_launch(const uint8_t *d_in, int w, int h, int stride_px, int mode, uint8_t *d_tmp, uint8_t *d_out,
                                       unsigned long long *d_acc, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CleanParams p{d_in, w, h, stride_px, d_tmp, d_out, d_acc};
  dim3 grid((w + 255) / 256, h);
  if (mode == 2) {
    premultiplied_kernel<<<grid, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
  }
  cudaError_t e = cudaMemsetAsync(d_acc, 0, 4 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return (int)e;
  dirty_sum_kernel<<<grid, 256, 0, st>>>(p);
  bleed_kernel<<<grid, 256, 0, st>>>(p);
  blur_kernel<<<grid, 256, 0, st>>>(p);
  return (int)cudaGetLastError();
}
