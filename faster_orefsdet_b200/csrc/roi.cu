// R1/P1: multi-level ROIAlign (NHWC).  R2+R3 (relation head) live in relation_tc.cu.
#include "common.cuh"

namespace fod {

// ------------------------------------------------------------------------------------------------
// ROIAlign, aligned=True, sampling_ratio=0 (adaptive grid), torchvision CPU semantics
// (torchvision/csrc/ops/cpu/roi_align_kernel.cpp + roi_align_common.h).
// One CTA per ROI; a warp owns a bin at a time; a lane owns 4 channels (16-byte loads, the
// four bilinear taps of a sample are four fully coalesced 512-byte rows of the NHWC map).
// ------------------------------------------------------------------------------------------------
struct RoiParams {
  const float* feat[FOD_MAX_LEVELS];
  int H[FOD_MAX_LEVELS], W[FOD_MAX_LEVELS], stride[FOD_MAX_LEVELS];
  int num_levels;
  int C;        // problems per image
  int roi_cap;
  int R;        // output resolution
  int tiled;    // output layout, see roi_align_kernel
  int pix;      // floats per pixel of the maps = channels (a multiple of 128; blockIdx.z selects the 128-channel block)
};

// d2 poolers.py:50-58, fp32 like torch: floor(4 + log2(sqrt(area)/224 + 1e-8)) clamped to the
// available levels.  min level = log2(stride[0]).
__device__ __forceinline__ int assign_level(float4 b, int min_level, int num_levels) {
  float area = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  float sz = __fsqrt_rn(area);
  float v = __fadd_rn(4.0f, log2f(__fadd_rn(__fdiv_rn(sz, 224.0f), 1e-8f)));
  v = floorf(v);
  v = fminf(fmaxf(v, (float)min_level), (float)(min_level + num_levels - 1));
  return (int)v - min_level;   // NaN area (negative) -> (int)NaN = 0 on CUDA; torch gives INT64_MIN: never valid input
}

// One bilinear sample coordinate along one axis (torchvision roi_align_common.h pre_calc_for_bilinear_interpolate):
// the two neighbouring rows / columns and their weights; w_low = w_high = 0 for a sample outside [-1, size].
struct AxisSample {
  int lo, hi;
  float w_hi, w_lo;  // weight of index `hi` (l) and of index `lo` (h = 1 - l)
};

__device__ __forceinline__ AxisSample axis_sample(float start, float bin, int p, int i, int grid, int size) {
  // coordinate = start + p*bin + (i+.5)*bin/grid     (left-to-right like the C++ expression)
  float v = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                      __fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), bin), (float)grid));
  AxisSample s;
  if (v < -1.0f || v > (float)size) {
    s.lo = s.hi = 0;
    s.w_hi = s.w_lo = 0.f;
    return s;
  }
  if (v <= 0.f) v = 0.f;
  int lo = (int)v, hi;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    v = (float)lo;
  } else {
    hi = lo + 1;
  }
  s.lo = lo;
  s.hi = hi;
  s.w_hi = __fsub_rn(v, (float)lo);
  s.w_lo = __fsub_rn(1.f, s.w_hi);
  return s;
}

constexpr int kMaxGrid = 7;             // sampling grid per bin handled by the shared-memory tables
constexpr int kMaxTaps = 8;             // distinct rows / columns one bin can touch (grid + 1)

// Per ROI and axis: for every bin the list of DISTINCT map rows (columns) its samples touch, with the summed
// interpolation weight of each.  Consecutive samples of a bin share a row (the upper neighbour of one is the lower
// neighbour of the next), so a bin with a g x g grid reads (g+1)^2 pixels instead of 4 g^2 taps.  Lists are padded
// to the ROI-wide maximum length with zero-weight entries so that the hot loop has a compile-time trip count.
constexpr int kMaxBins = 16;            // per axis (resolution 4, 8 or 14)
struct AxisTaps {
  int off[kMaxBins][kMaxTaps];   // element offset of the row / column inside the map (index * row pitch or * channels)
  float w[kMaxBins][kMaxTaps];
  int n[kMaxBins];
};

__device__ __forceinline__ void build_axis_taps(AxisTaps& t, int bin, float start, float bin_size, int grid, int size,
                                                int pitch) {
  int idx[kMaxTaps];
  float w[kMaxTaps];
  int n = 0;
  for (int i = 0; i < grid; ++i) {
    const AxisSample s = axis_sample(start, bin_size, bin, i, grid, size);
    if (s.w_lo == 0.f && s.w_hi == 0.f) continue;  // outside the map: contributes nothing
    // lower neighbour
    if (n > 0 && idx[n - 1] == s.lo) {
      w[n - 1] += s.w_lo;
    } else if (n > 1 && idx[n - 2] == s.lo) {
      w[n - 2] += s.w_lo;
    } else {
      idx[n] = s.lo;
      w[n] = s.w_lo;
      ++n;
    }
    // upper neighbour (lo == hi at the last row: both weights go to the same pixel, w_hi is 0 there)
    if (idx[n - 1] == s.hi) {
      w[n - 1] += s.w_hi;
    } else {
      idx[n] = s.hi;
      w[n] = s.w_hi;
      ++n;
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxTaps; ++k) {
    t.off[bin][k] = (k < n ? idx[k] : 0) * pitch;
    t.w[bin][k] = k < n ? w[k] : 0.f;
  }
  t.n[bin] = n;
}

struct RoiGeom {  // computed once per ROI by one thread
  const float* f;
  int W, nx, ny, tables, pad;
  float inv_count;
};

// NX = padded number of column taps per bin (compile time): the inner loop is 2 LDS.128 per bin and row plus
// LDG.128 + FMUL + 4 FFMA per tap.
template <int R, int NX>
__device__ __forceinline__ void roi_accumulate(const AxisTaps& xt, const AxisTaps& yt, const float* __restrict__ f, int ph,
                                               int pw0, int ny, float4 (&acc)[R / 2]) {
  for (int kr = 0; kr < ny; ++kr) {
    const float wy = yt.w[ph][kr];
    const float* rowp = f + yt.off[ph][kr];
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      int off[kMaxTaps];
      float w[kMaxTaps];
      *reinterpret_cast<int4*>(off) = *reinterpret_cast<const int4*>(&xt.off[pw0 + j][0]);
      *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(&xt.w[pw0 + j][0]);
      if (NX > 4) {
        *reinterpret_cast<int4*>(off + 4) = *reinterpret_cast<const int4*>(&xt.off[pw0 + j][4]);
        *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(&xt.w[pw0 + j][4]);
      }
      float4 v[NX];
#pragma unroll
      for (int k = 0; k < NX; ++k) v[k] = ldg4(rowp + off[k]);
#pragma unroll
      for (int k = 0; k < NX; ++k) {
        const float ww = wy * w[k];
        acc[j].x = fmaf(ww, v[k].x, acc[j].x);
        acc[j].y = fmaf(ww, v[k].y, acc[j].y);
        acc[j].z = fmaf(ww, v[k].z, acc[j].z);
        acc[j].w = fmaf(ww, v[k].w, acc[j].w);
      }
    }
  }
}

// One CTA per ROI; warp = (bin row, half of the bin columns), lane = 4 channels: every tap is one fully coalesced
// 512-byte row of the NHWC map, 4 accumulators per lane keep the register count low enough for full occupancy.
template <int R>
__global__ void __launch_bounds__(R * 64, R > 8 ? 1 : 2)
roi_align_kernel(RoiParams prm, const float* __restrict__ rois, const int32_t* __restrict__ roi_count,
                 float* __restrict__ pooled, int32_t* __restrict__ out_level) {
  __shared__ __align__(16) AxisTaps xt, yt;
  __shared__ RoiGeom geom;
  __shared__ float gbox[8];  // start_w, start_h, bin_w, bin_h, grid_w, grid_h (as float bits), W, H
  const int r = blockIdx.x, p = blockIdx.y;
  const int cnt = roi_count ? min(roi_count[p], prm.roi_cap) : prm.roi_cap;
  if (r >= cnt) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const float4 box = *reinterpret_cast<const float4*>(rois + ((size_t)p * prm.roi_cap + r) * 4);
    int min_level = 31 - __clz(prm.stride[0]);
    const int lvl = assign_level(box, min_level, prm.num_levels);
    if (out_level) out_level[(size_t)p * prm.roi_cap + r] = lvl;
    int H = prm.H[0], W = prm.W[0], stride = prm.stride[0];
    const float* fbase = prm.feat[0];
    if (lvl == 1) { H = prm.H[1]; W = prm.W[1]; stride = prm.stride[1]; fbase = prm.feat[1]; }
    if (lvl == 2) { H = prm.H[2]; W = prm.W[2]; stride = prm.stride[2]; fbase = prm.feat[2]; }
    const float scale = 1.0f / (float)stride;
    const float start_w = __fsub_rn(__fmul_rn(box.x, scale), 0.5f);
    const float start_h = __fsub_rn(__fmul_rn(box.y, scale), 0.5f);
    const float end_w = __fsub_rn(__fmul_rn(box.z, scale), 0.5f);
    const float end_h = __fsub_rn(__fmul_rn(box.w, scale), 0.5f);
    const float roi_w = __fsub_rn(end_w, start_w), roi_h = __fsub_rn(end_h, start_h);
    const int grid_h = (int)ceilf(__fdiv_rn(roi_h, (float)R));
    const int grid_w = (int)ceilf(__fdiv_rn(roi_w, (float)R));
    geom.f = fbase + (size_t)(p / prm.C) * H * W * prm.pix;
    geom.W = W;
    geom.inv_count = 1.0f / fmaxf((float)(grid_h * grid_w), 1.0f);
    geom.tables = grid_h <= kMaxGrid && grid_w <= kMaxGrid;
    gbox[0] = start_w;
    gbox[1] = start_h;
    gbox[2] = __fdiv_rn(roi_w, (float)R);
    gbox[3] = __fdiv_rn(roi_h, (float)R);
    gbox[4] = __int_as_float(grid_w);
    gbox[5] = __int_as_float(grid_h);
    gbox[6] = __int_as_float(W);
    gbox[7] = __int_as_float(H);
  }
  __syncthreads();
  const bool tables = geom.tables;
  const int W = __float_as_int(gbox[6]), H = __float_as_int(gbox[7]);
  const int grid_w = __float_as_int(gbox[4]), grid_h = __float_as_int(gbox[5]);
  if (tables) {
    if (threadIdx.x < R) build_axis_taps(xt, threadIdx.x, gbox[0], gbox[2], grid_w, W, prm.pix);
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + R) build_axis_taps(yt, threadIdx.x - 32, gbox[1], gbox[3], grid_h, H, W * prm.pix);
  }
  __syncthreads();
  const float* f = geom.f + blockIdx.z * kC + lane * 4;   // this CTA's 128-channel block of the pixel
  const int ph = warp >> 1, pw0 = (warp & 1) * (R / 2);
  float4 acc[R / 2];
#pragma unroll
  for (int j = 0; j < R / 2; ++j) acc[j] = make_float4(0, 0, 0, 0);
  if (tables) {
    int nx = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) nx = max(nx, xt.n[j]);
    const int ny = yt.n[ph];
    switch (nx) {
      case 0: break;
      case 1: case 2: roi_accumulate<R, 2>(xt, yt, f, ph, pw0, ny, acc); break;
      case 3: roi_accumulate<R, 3>(xt, yt, f, ph, pw0, ny, acc); break;
      case 4: roi_accumulate<R, 4>(xt, yt, f, ph, pw0, ny, acc); break;
      case 5: case 6: roi_accumulate<R, 6>(xt, yt, f, ph, pw0, ny, acc); break;
      default: roi_accumulate<R, 8>(xt, yt, f, ph, pw0, ny, acc); break;
    }
  } else {  // sampling grid beyond the tables (box much larger than the pyramid level expects): sample by sample
    for (int iy = 0; iy < grid_h; ++iy) {
      const AxisSample ys = axis_sample(gbox[1], gbox[3], ph, iy, grid_h, H);
      if (ys.w_lo == 0.f && ys.w_hi == 0.f) continue;
      const float* row_lo = f + (size_t)ys.lo * W * prm.pix;
      const float* row_hi = f + (size_t)ys.hi * W * prm.pix;
#pragma unroll
      for (int j = 0; j < R / 2; ++j) {
        for (int ix = 0; ix < grid_w; ++ix) {
          const AxisSample xs = axis_sample(gbox[0], gbox[2], pw0 + j, ix, grid_w, W);
          const float4 v1 = ldg4(row_lo + (size_t)xs.lo * prm.pix), v2 = ldg4(row_lo + (size_t)xs.hi * prm.pix);
          const float4 v3 = ldg4(row_hi + (size_t)xs.lo * prm.pix), v4 = ldg4(row_hi + (size_t)xs.hi * prm.pix);
          const float w1 = ys.w_lo * xs.w_lo, w2 = ys.w_lo * xs.w_hi, w3 = ys.w_hi * xs.w_lo, w4 = ys.w_hi * xs.w_hi;
          acc[j].x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
          acc[j].y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
          acc[j].z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
          acc[j].w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
        }
      }
    }
  }
  // Output layout.  tiled == 0: [P][roi_cap][R*R][128] (row-major ROI rows).  tiled == 1 (R == 8, consumed by
  // fod_relation_head): [P][units][256 k-chunks][128 rows][32], units = ceil(roi_cap / 128): the 16 KB A tile of
  // one 32-wide K chunk of 128 ROI rows is contiguous, so one TMA box fetches it as a linear stream.
  float* out;
  size_t bin_stride;
  if (prm.tiled) {
    const int units = (prm.roi_cap + 127) >> 7;
    out = pooled + ((((size_t)p * units + (r >> 7)) * 256 + (lane >> 3)) * 128 + (r & 127)) * 32 + (lane & 7) * 4;
    bin_stride = (size_t)4 * 128 * 32;  // next bin = 4 k-chunks further
  } else {
    out = pooled + ((size_t)p * prm.roi_cap + r) * R * R * prm.pix + blockIdx.z * kC + lane * 4;
    bin_stride = prm.pix;
  }
  const float inv_count = geom.inv_count;
#pragma unroll
  for (int j = 0; j < R / 2; ++j) {
    float4 o = make_float4(acc[j].x * inv_count, acc[j].y * inv_count, acc[j].z * inv_count, acc[j].w * inv_count);
    *reinterpret_cast<float4*>(out + (size_t)(ph * R + pw0 + j) * bin_stride) = o;
  }
}

}  // namespace fod

using namespace fod;

extern "C" int fod_roi_align_wide(const float* const* feat, const fod_level_t* levels, int num_levels, int batch,
                                  int problems_per_image, const float* rois, const int32_t* roi_count, int roi_cap,
                                  int resolution, int channels, int tiled, float* pooled, int32_t* out_level,
                                  fod_stream_t stream) {
  FOD_REQUIRE(feat && levels && rois && pooled, "fod_roi_align: null pointer");
  FOD_REQUIRE(num_levels >= 1 && num_levels <= FOD_MAX_LEVELS, "fod_roi_align: num_levels %d out of range", num_levels);
  FOD_REQUIRE(batch >= 0 && problems_per_image > 0 && roi_cap > 0, "fod_roi_align: bad sizes");
  FOD_REQUIRE(resolution == 8 || resolution == 4 || resolution == 14,
              "fod_roi_align: pooler resolution must be 8, 4 (POOLER_RESOLUTION / _2) or 14 (the C4 heads)");
  FOD_REQUIRE(channels >= kC && channels % kC == 0 && channels / kC <= 65535, "fod_roi_align: channels must be a multiple of 128");
  long P = (long)batch * problems_per_image;
  if (P == 0) return FOD_OK;
  FOD_REQUIRE(P <= 65535, "fod_roi_align: batch*classes %ld > 65535", P);
  RoiParams prm;
  for (int l = 0; l < num_levels; ++l) {
    FOD_REQUIRE(feat[l], "fod_roi_align: null level pointer");
    FOD_REQUIRE(levels[l].stride > 0 && (levels[l].stride & (levels[l].stride - 1)) == 0,
                "fod_roi_align: stride must be a power of two");
    if (l) FOD_REQUIRE(levels[l].stride == 2 * levels[l - 1].stride, "fod_roi_align: strides must form a pyramid");
    prm.feat[l] = feat[l];
    prm.H[l] = levels[l].height;
    prm.W[l] = levels[l].width;
    prm.stride[l] = levels[l].stride;
  }
  prm.num_levels = num_levels;
  prm.C = problems_per_image;
  prm.roi_cap = roi_cap;
  prm.R = resolution;
  prm.pix = channels;
  FOD_REQUIRE(!tiled || (resolution == 8 && channels == kC), "fod_roi_align: the tiled layout needs resolution 8 and 128 channels");
  prm.tiled = tiled ? 1 : 0;
  dim3 grid(roi_cap, (unsigned)P, (unsigned)(channels / kC));
  if (resolution == 8)
    roi_align_kernel<8><<<grid, 512, 0, as_stream(stream)>>>(prm, rois, roi_count, pooled, out_level);
  else if (resolution == 4)
    roi_align_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(prm, rois, roi_count, pooled, out_level);
  else
    roi_align_kernel<14><<<grid, 896, 0, as_stream(stream)>>>(prm, rois, roi_count, pooled, out_level);
  FOD_CUDA_LAUNCH_CHECK("fod_roi_align");
  return FOD_OK;
}

extern "C" int fod_roi_align(const float* const* feat, const fod_level_t* levels, int num_levels, int batch,
                             int problems_per_image, const float* rois, const int32_t* roi_count, int roi_cap,
                             int resolution, int tiled, float* pooled, int32_t* out_level, fod_stream_t stream) {
  return fod_roi_align_wide(feat, levels, num_levels, batch, problems_per_image, rois, roi_count, roi_cap, resolution, kC,
                            tiled, pooled, out_level, stream);
}
