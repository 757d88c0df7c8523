// GroupNorm (+ ReLU) over NHWC fp32 maps: the normalisation between the tower convolution and agn_hm / bbox_pred of
// CenterNetHead (CenterNet2/centernet/modeling/dense_heads/centernet_head.py:61-72, 145-150; GroupNorm(32, 128)).
// ATen's group_norm returns an NCHW tensor for an NHWC input (two layout copies) and cannot fuse the ReLU.
// Two memory-bound passes: (1) per-(map, group) sum / sum of squares, fp32 partials per thread, fp64 across threads
// and CTAs; (2) y = relu((x - mean) * rstd * gamma + beta), one float4 (four channels of one group) per thread.
// Algorithmic bytes: 2 reads + 1 write of the map (the second read hits L2 when the map fits).
#include "common.cuh"

namespace fod {
namespace gn {

constexpr int kThreads = 256;

// grid (slabs, maps); channels % 4 == 0, channels-per-group % 4 == 0, channels / 4 <= kThreads
__global__ void __launch_bounds__(kThreads) stats_kernel(const float* __restrict__ x, long hw, int channels, int cpg,
                                                         int groups, double* __restrict__ stats) {
  extern __shared__ double red[];  // [groups][2]
  const int c4n = channels >> 2;
  const int c4 = threadIdx.x % c4n, prow = threadIdx.x / c4n, rows = kThreads / c4n;
  for (int i = threadIdx.x; i < 2 * groups; i += kThreads) red[i] = 0.0;
  __syncthreads();
  const long per = (hw + gridDim.x - 1) / gridDim.x;
  const long p0 = (long)blockIdx.x * per, p1 = p0 + per < hw ? p0 + per : hw;
  const float* base = x + ((size_t)blockIdx.y * hw) * channels + c4 * 4;
  float s = 0.f, ss = 0.f;
  if (prow < rows)
#pragma unroll 4
    for (long p = p0 + prow; p < p1; p += rows) {
      const float4 v = ldg4(base + (size_t)p * channels);
      s += (v.x + v.y) + (v.z + v.w);
      ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
    }
  if (prow < rows) {
    const int g = (c4 * 4) / cpg;
    atomicAdd(&red[2 * g], (double)s);
    atomicAdd(&red[2 * g + 1], (double)ss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * groups; i += kThreads)
    atomicAdd(&stats[(size_t)blockIdx.y * 2 * groups + i], red[i]);
}

__global__ void __launch_bounds__(kThreads) apply_kernel(const float* __restrict__ x, float* __restrict__ y, long hw,
                                                         int channels, int cpg, int groups, const double* __restrict__ stats,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, int relu, size_t total4, float* __restrict__ y_amax,
                                                         int amax_per_map) {
  // grid = (CTAs per map, maps); total4 = float4 elements of ONE map
  const int c4n = channels >> 2;
  float vmax = 0.f;
  const double inv_n = 1.0 / ((double)hw * cpg);
  const size_t map = blockIdx.y;
  for (size_t il = (size_t)blockIdx.x * kThreads + threadIdx.x; il < total4; il += (size_t)gridDim.x * kThreads) {
    const size_t i = map * total4 + il;
    const int c4 = (int)(il % c4n);
    const int g = (c4 * 4) / cpg;
    const double* st = stats + (map * groups + g) * 2;
    const double mean = st[0] * inv_n;
    double var = st[1] * inv_n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float mu = (float)mean;
    const float4 v = ldg4(x + i * 4);
    const float4 ga = gamma ? ldg4(gamma + c4 * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 be = beta ? ldg4(beta + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = (v.x - mu) * rstd * ga.x + be.x;
    o.y = (v.y - mu) * rstd * ga.y + be.y;
    o.z = (v.z - mu) * rstd * ga.z + be.z;
    o.w = (v.w - mu) * rstd * ga.w + be.w;
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    *reinterpret_cast<float4*>(y + i * 4) = o;
    vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
  }
  if (y_amax) {
    const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
    if ((threadIdx.x & 31) == 0 && w) atomicMax(reinterpret_cast<unsigned int*>(y_amax + (amax_per_map ? map : 0)), w);
  }
}

// GroupNorm as a per-(map, channel) affine map from the per-tile channel sums / sums of squares that the producing
// convolution wrote (fod_conv2d_nhwc colsum / colsumsq): scale = rstd * gamma, shift = beta - mean * scale.  The consumer
// applies x * scale + shift (+ ReLU) to its input operand, so the normalised map never exists in memory.
// One CTA per map; fp64 across tiles and channels of a group; bound = max_c(|scale_c| * x_amax + |shift_c|).
__global__ void __launch_bounds__(256) affine_kernel(const float* __restrict__ colsum, const float* __restrict__ colsumsq,
                                                     int tiles, int channels, int cpg, double inv_n,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                     const float* __restrict__ x_amax, float* __restrict__ scale,
                                                     float* __restrict__ shift, float* __restrict__ y_amax, int amax_per_map) {
  extern __shared__ double sm[];   // [channels] sums, [channels] sums of squares
  double* s1 = sm;
  double* s2 = sm + channels;
  const size_t base = (size_t)blockIdx.x * tiles * channels;
  for (int c = threadIdx.x; c < channels; c += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int t = 0; t < tiles; ++t) {
      a += (double)colsum[base + (size_t)t * channels + c];
      b += (double)colsumsq[base + (size_t)t * channels + c];
    }
    s1[c] = a;
    s2[c] = b;
  }
  __syncthreads();
  float vmax = 0.f;
  const float xa = x_amax ? __ldg(x_amax + (amax_per_map ? blockIdx.x : 0)) : 0.f;
  for (int c = threadIdx.x; c < channels; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    double a = 0.0, b = 0.0;
    for (int k = 0; k < cpg; ++k) {
      a += s1[g0 + k];
      b += s2[g0 + k];
    }
    const double mean = a * inv_n;
    double var = b * inv_n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = rstd * (gamma ? gamma[c] : 1.f);
    const float sh = (beta ? beta[c] : 0.f) - (float)mean * sc;
    scale[(size_t)blockIdx.x * channels + c] = sc;
    shift[(size_t)blockIdx.x * channels + c] = sh;
    vmax = fmaxf(vmax, fabsf(sc) * xa + fabsf(sh));
  }
  if (y_amax) {
    const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
    if ((threadIdx.x & 31) == 0 && w) atomicMax(reinterpret_cast<unsigned int*>(y_amax + (amax_per_map ? blockIdx.x : 0)), w);
  }
}

}  // namespace gn
}  // namespace fod

using namespace fod;

extern "C" size_t fod_group_norm_workspace_bytes(int maps, int groups) { return (size_t)maps * groups * 2 * sizeof(double); }

extern "C" int fod_group_norm_nhwc(const float* x, int maps, long hw, int channels, int groups, const float* gamma,
                                   const float* beta, float eps, int relu, float* y, float* y_amax, int amax_per_map,
                                   void* workspace, fod_stream_t stream) {
  FOD_REQUIRE(x && y && workspace, "fod_group_norm_nhwc: null pointer");
  FOD_REQUIRE(maps >= 0 && hw > 0 && channels > 0 && groups > 0 && channels % groups == 0, "fod_group_norm_nhwc: bad sizes");
  const int cpg = channels / groups;
  FOD_REQUIRE(channels % 4 == 0 && cpg % 4 == 0 && channels / 4 <= gn::kThreads,
              "fod_group_norm_nhwc: channels per group must be a multiple of 4 and channels <= %d", 4 * gn::kThreads);
  FOD_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)workspace) & 15) == 0,
              "fod_group_norm_nhwc: pointers must be 16-byte aligned");
  if (maps == 0) return FOD_OK;
  FOD_CUDA_CALL(cudaMemsetAsync(workspace, 0, fod_group_norm_workspace_bytes(maps, groups), as_stream(stream)));
  // Slabs of 64 pixels, whatever the number of maps: the per-thread fp32 partial sums of a map are then the same in
  // every batch (the statistics of a map must not depend on its batch mates); 100 slabs per 80 x 80 map keep the pass at
  // several CTAs per SM (it is bound by the latency of its loads).
  long slabs = (hw + 63) / 64;
  if (slabs < 1) slabs = 1;
  gn::stats_kernel<<<dim3((unsigned)slabs, (unsigned)maps), gn::kThreads, 2 * groups * sizeof(double), as_stream(stream)>>>(
      x, hw, channels, cpg, groups, static_cast<double*>(workspace));
  FOD_CUDA_LAUNCH_CHECK("fod_group_norm_nhwc (stats)");
  const size_t total4 = (size_t)hw * (channels / 4);      // per map
  size_t blocks = (total4 + gn::kThreads - 1) / gn::kThreads;
  const size_t want = (size_t)(148 * 16 + maps - 1) / maps;
  if (blocks > want) blocks = want;
  if (blocks < 1) blocks = 1;
  gn::apply_kernel<<<dim3((unsigned)blocks, (unsigned)maps), gn::kThreads, 0, as_stream(stream)>>>(
      x, y, hw, channels, cpg, groups, static_cast<const double*>(workspace), gamma, beta, eps, relu, total4, y_amax,
      amax_per_map);
  FOD_CUDA_LAUNCH_CHECK("fod_group_norm_nhwc (apply)");
  return FOD_OK;
}

extern "C" int fod_group_norm_affine(const float* colsum, const float* colsumsq, int maps, int tiles_per_map, int channels,
                                     int groups, long hw, const float* gamma, const float* beta, float eps,
                                     const float* x_amax, float* scale, float* shift, float* y_amax, int amax_per_map,
                                     fod_stream_t stream) {
  FOD_REQUIRE(colsum && colsumsq && scale && shift, "fod_group_norm_affine: null pointer");
  FOD_REQUIRE(maps >= 0 && tiles_per_map > 0 && channels > 0 && groups > 0 && channels % groups == 0 && hw > 0 &&
                  channels <= 2048, "fod_group_norm_affine: bad sizes");
  if (maps == 0) return FOD_OK;
  const int cpg = channels / groups;
  gn::affine_kernel<<<maps, 256, 2 * channels * sizeof(double), as_stream(stream)>>>(
      colsum, colsumsq, tiles_per_map, channels, cpg, 1.0 / ((double)hw * cpg), gamma, beta, eps, x_amax, scale, shift, y_amax, amax_per_map);
  FOD_CUDA_LAUNCH_CHECK("fod_group_norm_affine");
  return FOD_OK;
}
