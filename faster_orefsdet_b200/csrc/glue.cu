// Memory-bound glue between the tensor-core convolutions of the feature extractor (d2!/modeling/backbone/vovnet.py):
//   * fod_stem_patches   - im2col of the 3-channel stem convolution (3x3, stride 2, pad 1, vovnet.py stem_1): each output
//                          pixel gets its 27 input values (+5 zeros) as one 128-byte row, so stem_1 runs as a 1x1
//                          tensor-core convolution with K = 32 instead of a 3-channel cuDNN kernel.
//   * fod_maxpool3x3s2   - nn.MaxPool2d(3, stride 2, ceil_mode=True) of the OSA stages (vovnet.py _OSA_stage) over NHWC
//                          maps, optionally multiplied by the eSE gate of the producing stage (x * hsigmoid(fc(avg(x))),
//                          vovnet.py eSEModule; the gate is >= 0, so it commutes with the max), written straight into a
//                          channel slice of the next stage's concat buffer.
// One float4 (4 channels) per thread, coalesced 128-bit accesses; the 9-fold window overlap is served by L1/L2.
#include "common.cuh"

namespace fod {
namespace glue {

constexpr int kThreads = 256;

// x: [N][H][W][3] fp32 (normalised image, NHWC) -> p: [N][Ho][Wo][32], k = (ky*3 + kx)*3 + c, rows 27..31 zero
__global__ void __launch_bounds__(kThreads) stem_patches_kernel(const float* __restrict__ x, int H, int W, int Ho, int Wo,
                                                                float* __restrict__ p, size_t total) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    const int q = (int)(i & 7);  // float4 index inside the 32-wide row
    const size_t pix = i >> 3;
    const int ox = (int)(pix % Wo);
    const size_t r = pix / Wo;
    const int oy = (int)(r % Ho);
    const size_t n = r / Ho;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = q * 4 + j;
      float val = 0.f;
      if (k < 27) {
        const int tap = k / 3, c = k - tap * 3;
        const int ky = tap / 3, kx = tap - ky * 3;
        const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = __ldg(x + ((n * H + iy) * (size_t)W + ix) * 3 + c);
      }
      v[j] = val;
    }
    *reinterpret_cast<float4*>(p + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Same rows straight from the raw image: x [N][3][H][W] uint8 planar (the dataset mapper's CHW tensors, stacked),
// normalised on the fly as (x - mean[c]) / std[c] (fsod_cen.py:540-555; IEEE division, identical to the torch ops).
__global__ void __launch_bounds__(kThreads) stem_patches_u8_kernel(const uint8_t* __restrict__ x, int H, int W, int Ho, int Wo,
                                                                   float m0, float m1, float m2, float s0, float s1, float s2,
                                                                   float* __restrict__ p, size_t total) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    const int q = (int)(i & 7);
    const size_t pix = i >> 3;
    const int ox = (int)(pix % Wo);
    const size_t r = pix / Wo;
    const int oy = (int)(r % Ho);
    const size_t n = r / Ho;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = q * 4 + j;
      float val = 0.f;
      if (k < 27) {
        const int tap = k / 3, c = k - tap * 3;
        const int ky = tap / 3, kx = tap - ky * 3;
        const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const float raw = (float)__ldg(x + ((n * 3 + c) * (size_t)H + iy) * W + ix);
          const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
          val = __fdiv_rn(__fsub_rn(raw, mean), sd);
        }
      }
      v[j] = val;
    }
    *reinterpret_cast<float4*>(p + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// stem_1 in one pass on the CUDA cores: raw uint8 planar image -> normalise -> 3x3 / stride 2 / pad 1 convolution with 64
// output channels (BN folded) -> ReLU -> NHWC fp32.  27 x 64 FMAs per output pixel are too few for the tensor-core
// pipeline to pay (the im2col rows alone are 2.5x the bytes of the result), and the layer is bound by writing its
// 1.7 GB result.  Four adjacent lanes share a pixel; lane j owns channels {16 i + 4 j .. + 3 : i = 0..3}, so every
// store instruction of a warp writes 8 x 64 contiguous bytes; the weights sit in shared memory as [27][64].
constexpr int kStemC = 64;
__global__ void __launch_bounds__(kThreads) stem1_u8_kernel(const uint8_t* __restrict__ x, int H, int W, int Ho, int Wo,
                                                            float m0, float m1, float m2, float s0, float s1, float s2,
                                                            const float* __restrict__ w /*[64][27] = OIHW*/,
                                                            const float* __restrict__ bias, float* __restrict__ y,
                                                            float* __restrict__ y_amax, size_t total_pairs, int pairs_per_row) {
  __shared__ __align__(16) float ws[27][kStemC];
  __shared__ __align__(16) float bs[kStemC];
  __shared__ float lut[3][257];   // (v - mean[c]) / std[c] for the 256 raw values (IEEE division, as torch); [256] = padding
  for (int i = threadIdx.x; i < 27 * kStemC; i += kThreads) {
    const int co = i % kStemC, k = i / kStemC;         // k = (ky*3 + kx)*3 + c  <-  OIHW index (c*3 + ky)*3 + kx
    const int tap = k / 3, c = k - tap * 3;
    ws[k][co] = w[co * 27 + c * 9 + tap];
  }
  for (int i = threadIdx.x; i < kStemC; i += kThreads) bs[i] = bias ? bias[i] : 0.f;
  for (int i = threadIdx.x; i < 3 * 257; i += kThreads) {
    const int c = i / 257, v = i - c * 257;
    lut[c][v] = v == 256 ? 0.f : __fdiv_rn(__fsub_rn((float)v, c == 0 ? m0 : (c == 1 ? m1 : m2)), c == 0 ? s0 : (c == 1 ? s1 : s2));
  }
  __syncthreads();
  const int j = threadIdx.x & 3;
  const size_t plane = (size_t)H * W;
  float vmax = 0.f;
  // a group of four lanes computes two horizontally adjacent output pixels (their windows share an input column)
  for (size_t pr = ((size_t)blockIdx.x * kThreads + threadIdx.x) >> 2; pr < total_pairs; pr += ((size_t)gridDim.x * kThreads) >> 2) {
    const int px = (int)(pr % pairs_per_row);
    const size_t r = pr / pairs_per_row;
    const int oy = (int)(r % Ho);
    const size_t n = r / Ho;
    const int ox = 2 * px;
    const bool second = ox + 1 < Wo;
    float4 acc[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[0][i] = acc[1][i] = *reinterpret_cast<const float4*>(&bs[16 * i + 4 * j]);
    const uint8_t* img = x + n * 3 * plane;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy + ky - 1;
      const bool rowin = iy >= 0 && iy < H;
      const uint8_t* rowp = img + (size_t)(rowin ? iy : 0) * W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v[5];   // input columns 2*ox - 1 .. 2*ox + 3 of this row and channel
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int ix = 2 * ox - 1 + q;
          const bool in = rowin && ix >= 0 && ix < W;
          const int raw = in ? (int)__ldg(rowp + c * plane + ix) : 256;
          v[q] = lut[c][raw];
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int k = (ky * 3 + kx) * 3 + c;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 wv = *reinterpret_cast<const float4*>(&ws[k][16 * i + 4 * j]);
            acc[0][i].x = fmaf(v[kx], wv.x, acc[0][i].x); acc[0][i].y = fmaf(v[kx], wv.y, acc[0][i].y);
            acc[0][i].z = fmaf(v[kx], wv.z, acc[0][i].z); acc[0][i].w = fmaf(v[kx], wv.w, acc[0][i].w);
            acc[1][i].x = fmaf(v[kx + 2], wv.x, acc[1][i].x); acc[1][i].y = fmaf(v[kx + 2], wv.y, acc[1][i].y);
            acc[1][i].z = fmaf(v[kx + 2], wv.z, acc[1][i].z); acc[1][i].w = fmaf(v[kx + 2], wv.w, acc[1][i].w);
          }
        }
      }
    }
    const size_t pix = (n * Ho + oy) * (size_t)Wo + ox;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (p == 1 && !second) break;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = acc[p][i];
        const float4 o = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
        vmax = fmaxf(fmaxf(vmax, fmaxf(o.x, o.y)), fmaxf(o.z, o.w));
        *reinterpret_cast<float4*>(y + (pix + p) * kStemC + 16 * i + 4 * j) = o;
      }
    }
  }
  if (y_amax) {
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
    if ((threadIdx.x & 31) == 0 && wm) atomicMax(reinterpret_cast<unsigned int*>(y_amax), wm);
  }
}

// x: [N][H][W] pixels of xs floats (first C used), gate: [N][C] or null -> y: [N][Ho][Wo] pixels of ys floats
__global__ void __launch_bounds__(kThreads) maxpool_kernel(const float* __restrict__ x, long xs, int H, int W, int C,
                                                           const float* __restrict__ gate, float* __restrict__ y, long ys,
                                                           int Ho, int Wo, size_t total) {
  const int c4n = C >> 2;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    const int c4 = (int)(i % c4n);
    const size_t pix = i / c4n;
    const int ox = (int)(pix % Wo);
    const size_t r = pix / Wo;
    const int oy = (int)(r % Ho);
    const size_t n = r / Ho;
    const int y0 = 2 * oy, x0 = 2 * ox;
    const int y1 = min(y0 + 3, H), x1 = min(x0 + 3, W);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int iy = y0; iy < y1; ++iy)
      for (int ix = x0; ix < x1; ++ix) {
        const float4 v = ldg4(x + ((n * H + iy) * (size_t)W + ix) * xs + c4 * 4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    if (gate) {
      const float4 g = ldg4(gate + n * C + c4 * 4);
      m.x *= g.x; m.y *= g.y; m.z *= g.z; m.w *= g.w;
    }
    *reinterpret_cast<float4*>(y + pix * ys + c4 * 4) = m;
  }
}

}  // namespace glue
}  // namespace fod

using namespace fod;

static unsigned grid_for(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = 148 * 32;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

extern "C" int fod_stem_patches(const float* x, int n, int h, int w, float* patches, fod_stream_t stream) {
  FOD_REQUIRE(x && patches, "fod_stem_patches: null pointer");
  FOD_REQUIRE(n >= 0 && h > 0 && w > 0, "fod_stem_patches: bad sizes");
  FOD_REQUIRE(((uintptr_t)patches & 15) == 0, "fod_stem_patches: output must be 16-byte aligned");
  if (n == 0) return FOD_OK;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const size_t total = (size_t)n * ho * wo * 8;
  glue::stem_patches_kernel<<<grid_for(total, glue::kThreads), glue::kThreads, 0, as_stream(stream)>>>(x, h, w, ho, wo, patches,
                                                                                                    total);
  FOD_CUDA_LAUNCH_CHECK("fod_stem_patches");
  return FOD_OK;
}

extern "C" int fod_stem_patches_u8(const uint8_t* x, int n, int h, int w, const float* mean3, const float* std3, float* patches,
                                   fod_stream_t stream) {
  FOD_REQUIRE(x && patches && mean3 && std3, "fod_stem_patches_u8: null pointer");
  FOD_REQUIRE(n >= 0 && h > 0 && w > 0, "fod_stem_patches_u8: bad sizes");
  FOD_REQUIRE(((uintptr_t)patches & 15) == 0, "fod_stem_patches_u8: output must be 16-byte aligned");
  if (n == 0) return FOD_OK;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const size_t total = (size_t)n * ho * wo * 8;
  glue::stem_patches_u8_kernel<<<grid_for(total, glue::kThreads), glue::kThreads, 0, as_stream(stream)>>>(
      x, h, w, ho, wo, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], patches, total);
  FOD_CUDA_LAUNCH_CHECK("fod_stem_patches_u8");
  return FOD_OK;
}

extern "C" int fod_stem1_u8(const uint8_t* x, int n, int h, int w, const float* mean3, const float* std3, const float* weight,
                            const float* bias, float* y, float* y_amax, fod_stream_t stream) {
  FOD_REQUIRE(x && y && mean3 && std3 && weight, "fod_stem1_u8: null pointer");
  FOD_REQUIRE(n >= 0 && h > 0 && w > 0, "fod_stem1_u8: bad sizes");
  FOD_REQUIRE(((uintptr_t)y & 15) == 0, "fod_stem1_u8: output must be 16-byte aligned");
  if (n == 0) return FOD_OK;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const int pairs_per_row = (wo + 1) / 2;
  const size_t total_pairs = (size_t)n * ho * pairs_per_row;
  glue::stem1_u8_kernel<<<grid_for(total_pairs * 4, glue::kThreads), glue::kThreads, 0, as_stream(stream)>>>(
      x, h, w, ho, wo, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], weight, bias, y, y_amax, total_pairs,
      pairs_per_row);
  FOD_CUDA_LAUNCH_CHECK("fod_stem1_u8");
  return FOD_OK;
}

extern "C" int fod_maxpool3x3s2_nhwc(const float* x, int n, int h, int w, int c, long x_pixel_stride, const float* gate,
                                     float* y, long y_pixel_stride, fod_stream_t stream) {
  FOD_REQUIRE(x && y, "fod_maxpool3x3s2_nhwc: null pointer");
  FOD_REQUIRE(n >= 0 && h >= 3 && w >= 3 && c > 0 && c % 4 == 0 && x_pixel_stride % 4 == 0 && y_pixel_stride % 4 == 0 &&
                  x_pixel_stride >= c && y_pixel_stride >= c, "fod_maxpool3x3s2_nhwc: bad sizes (channels / strides multiples of 4)");
  FOD_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gate) & 15) == 0, "fod_maxpool3x3s2_nhwc: pointers must be 16-byte aligned");
  if (n == 0) return FOD_OK;
  // ceil_mode output size of nn.MaxPool2d(3, 2): ceil((H - 3) / 2) + 1 (the last window always starts inside the map)
  const int ho = (h - 3 + 1) / 2 + 1, wo = (w - 3 + 1) / 2 + 1;
  const size_t total = (size_t)n * ho * wo * (c / 4);
  glue::maxpool_kernel<<<grid_for(total, glue::kThreads), glue::kThreads, 0, as_stream(stream)>>>(
      x, x_pixel_stride, h, w, c, gate, y, y_pixel_stride, ho, wo, total);
  FOD_CUDA_LAUNCH_CHECK("fod_maxpool3x3s2_nhwc");
  return FOD_OK;
}
