// Q1 (support taps) and Q2+Q3 (depthwise correlation fused with the 1x1 relation conv).
//
// fod_correlate, v1 (fp32 CUDA-core contraction):
//   tile = 64 consecutive pixels x 128 output channels per CTA, 256 threads,
//   4x8 register tile per thread.  The A operand [64 px][256 k] = [s | q] is never
//   materialised in HBM: per 16-channel chunk each thread rebuilds s for one pixel
//   and four channels from the 3x3 neighbourhood of q (L1/L2 resident; NHWC so every
//   tap is a 16-byte load, a quarter-warp covers one pixel's chunk) and drops s and q
//   k-major into shared memory next to the matching two 128x16 slices of W3.
//   HBM traffic = read q once + write attn once (1024 B per pixel per problem).
#include "common.cuh"

namespace fod {

// ------------------------------------------------------------------------------------------------
// Q1: adaptive average pools (1,1), (1,3), (3,1) of the prototype.
// PyTorch adaptive pooling bins: start = floor(i*n/out), end = ceil((i+1)*n/out).
// ------------------------------------------------------------------------------------------------
__global__ void support_taps_kernel(const float* __restrict__ proto, int h, int w, float* __restrict__ taps) {
  const int c = blockIdx.x;          // class
  const int ch = threadIdx.x;        // channel 0..127
  const float* P = proto + (size_t)c * h * w * kC;
  float col[3] = {0.f, 0.f, 0.f}, row[3] = {0.f, 0.f, 0.f}, all = 0.f;
  int cs[3], ce[3], rs[3], re[3];
  for (int i = 0; i < 3; ++i) {
    cs[i] = (i * w) / 3;
    ce[i] = ((i + 1) * w + 2) / 3;
    rs[i] = (i * h) / 3;
    re[i] = ((i + 1) * h + 2) / 3;
  }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float v = P[((size_t)y * w + x) * kC + ch];
      all += v;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (x >= cs[i] && x < ce[i]) col[i] += v;
        if (y >= rs[i] && y < re[i]) row[i] += v;
      }
    }
  float* T = taps + (size_t)c * 7 * kC;
  T[0 * kC + ch] = all / (float)(h * w);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T[(1 + i) * kC + ch] = col[i] / (float)(h * (ce[i] - cs[i]));
    T[(4 + i) * kC + ch] = row[i] / (float)(w * (re[i] - rs[i]));
  }
}

// ------------------------------------------------------------------------------------------------
// Q2+Q3
// ------------------------------------------------------------------------------------------------
constexpr int kBM = 64;    // pixels per CTA
constexpr int kBN = 128;   // output channels
constexpr int kBK = 16;    // channels per chunk
constexpr int kCorrThreads = 256;
constexpr int kAPad = 4, kWPad = 4;

__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__global__ void __launch_bounds__(kCorrThreads, 2)
correlate_kernel(const float* __restrict__ q, const float* __restrict__ taps, const float* __restrict__ w3,
                 const float* __restrict__ b3, float* __restrict__ attn, int C, int H, int W) {
  __shared__ __align__(16) float As[2][kBK][kBM + kAPad];   // [0] = s chunk, [1] = q chunk, k-major
  __shared__ __align__(16) float Ws[2][kBK][kBN + kWPad];   // matching W3 slices, k-major
  __shared__ __align__(16) float Ts[7][kC];
  const int tid = threadIdx.x;
  const int p = blockIdx.y, b = p / C, c = p - b * C;
  const int HW = H * W;
  const int px0 = blockIdx.x * kBM;
  const float* qb = q + (size_t)b * HW * kC;
  for (int i = tid; i < 7 * kC; i += kCorrThreads) (&Ts[0][0])[i] = taps[(size_t)c * 7 * kC + i];
  // loader mapping: one pixel row of the tile, 4 channels of the chunk
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int px = px0 + lr;
  const bool pvalid = px < HW;
  const int y = pvalid ? px / W : 0, x = pvalid ? px - y * W : 0;
  // compute mapping
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  __syncthreads();
  for (int ch0 = 0; ch0 < kC; ch0 += kBK) {
    const int ch = ch0 + lk;
    float4 sv = make_float4(0, 0, 0, 0), qc = make_float4(0, 0, 0, 0);
    if (pvalid) {
      const float4 k11 = *reinterpret_cast<const float4*>(&Ts[0][ch]);
      const float4 k13l = *reinterpret_cast<const float4*>(&Ts[1][ch]);
      const float4 k13c = *reinterpret_cast<const float4*>(&Ts[2][ch]);
      const float4 k13r = *reinterpret_cast<const float4*>(&Ts[3][ch]);
      float4 hrow[3];
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        int yy = y + dy;
        float4 hv = make_float4(0, 0, 0, 0);
        if (yy >= 0 && yy < H) {
          const float* rowp = qb + ((size_t)yy * W + x) * kC + ch;
          float4 m = ldg4(rowp);
          if (dy == 0) qc = m;
          hv = mul4(k13c, m);
          if (x > 0) hv = fma4(k13l, ldg4(rowp - kC), hv);
          if (x + 1 < W) hv = fma4(k13r, ldg4(rowp + kC), hv);
          hv = relu4(hv);
        }
        hrow[dy + 1] = hv;
      }
      const float4 k31u = *reinterpret_cast<const float4*>(&Ts[4][ch]);
      const float4 k31c = *reinterpret_cast<const float4*>(&Ts[5][ch]);
      const float4 k31d = *reinterpret_cast<const float4*>(&Ts[6][ch]);
      float4 bv = relu4(fma4(k31d, hrow[2], fma4(k31c, hrow[1], mul4(k31u, hrow[0]))));
      float4 av = relu4(mul4(k11, relu4(mul4(k11, qc))));
      sv = add4(add4(av, bv), qc);
    }
    // W3 slices for this chunk: rows n = lr and lr+64, columns ch (s half) and 128+ch (q half)
    float4 wa0 = ldg4(w3 + (size_t)lr * 256 + ch), wa1 = ldg4(w3 + (size_t)(lr + 64) * 256 + ch);
    float4 wq0 = ldg4(w3 + (size_t)lr * 256 + 128 + ch), wq1 = ldg4(w3 + (size_t)(lr + 64) * 256 + 128 + ch);
    __syncthreads();  // previous chunk fully consumed
    As[0][lk + 0][lr] = sv.x; As[0][lk + 1][lr] = sv.y; As[0][lk + 2][lr] = sv.z; As[0][lk + 3][lr] = sv.w;
    As[1][lk + 0][lr] = qc.x; As[1][lk + 1][lr] = qc.y; As[1][lk + 2][lr] = qc.z; As[1][lk + 3][lr] = qc.w;
    Ws[0][lk + 0][lr] = wa0.x; Ws[0][lk + 1][lr] = wa0.y; Ws[0][lk + 2][lr] = wa0.z; Ws[0][lk + 3][lr] = wa0.w;
    Ws[0][lk + 0][lr + 64] = wa1.x; Ws[0][lk + 1][lr + 64] = wa1.y; Ws[0][lk + 2][lr + 64] = wa1.z; Ws[0][lk + 3][lr + 64] = wa1.w;
    Ws[1][lk + 0][lr] = wq0.x; Ws[1][lk + 1][lr] = wq0.y; Ws[1][lk + 2][lr] = wq0.z; Ws[1][lk + 3][lr] = wq0.w;
    Ws[1][lk + 0][lr + 64] = wq1.x; Ws[1][lk + 1][lr + 64] = wq1.y; Ws[1][lk + 2][lr + 64] = wq1.z; Ws[1][lk + 3][lr + 64] = wq1.w;
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int k = 0; k < kBK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[half][k][ty * 4]);
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[half][k][tx * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[half][k][tx * 8 + 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
    }
  }
  // epilogue: bias + ReLU, NHWC store
  const float4 bb0 = ldg4(b3 + tx * 8), bb1 = ldg4(b3 + tx * 8 + 4);
  const float bv[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
  float* ob = attn + (size_t)p * HW * kC;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int opx = px0 + ty * 4 + i;
    if (opx < HW) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(acc[i][j] + bv[j], 0.f);
      float4* dst = reinterpret_cast<float4*>(ob + (size_t)opx * kC + tx * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

}  // namespace fod

using namespace fod;

extern "C" int fod_support_taps(const float* proto, int num_classes, int h, int w, float* taps, fod_stream_t stream) {
  FOD_REQUIRE(proto && taps, "fod_support_taps: null pointer");
  FOD_REQUIRE(num_classes >= 0 && h > 0 && w > 0, "fod_support_taps: bad sizes");
  if (num_classes == 0) return FOD_OK;
  support_taps_kernel<<<num_classes, kC, 0, as_stream(stream)>>>(proto, h, w, taps);
  FOD_CUDA_LAUNCH_CHECK("fod_support_taps");
  return FOD_OK;
}

extern "C" int fod_correlate(const float* q, const float* taps, const float* w3, const float* b3, float* attn, int batch,
                             int num_classes, int height, int width, fod_stream_t stream) {
  FOD_REQUIRE(q && taps && w3 && b3 && attn, "fod_correlate: null pointer");
  FOD_REQUIRE(batch >= 0 && num_classes >= 0 && height > 0 && width > 0, "fod_correlate: bad sizes");
  FOD_REQUIRE((((uintptr_t)q | (uintptr_t)attn | (uintptr_t)w3 | (uintptr_t)b3) & 15) == 0,
              "fod_correlate: pointers must be 16-byte aligned");
  long P = (long)batch * num_classes;
  if (P == 0) return FOD_OK;
  FOD_REQUIRE(P <= 65535, "fod_correlate: batch*classes %ld > 65535", P);
  dim3 grid((height * width + kBM - 1) / kBM, (unsigned)P);
  correlate_kernel<<<grid, kCorrThreads, 0, as_stream(stream)>>>(q, taps, w3, b3, attn, num_classes, height, width);
  FOD_CUDA_LAUNCH_CHECK("fod_correlate");
  return FOD_OK;
}
