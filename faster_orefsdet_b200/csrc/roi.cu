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
#ifdef FOD_ROI_PROF
__device__ long long g_roi_prof[8];
#define RPROF(slot) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 64) atomicAdd((unsigned long long*)&g_roi_prof[slot], (unsigned long long)(clock64() - t_start)); } while (0)
#else
#define RPROF(slot)
#endif

// Everything the pooling of one ROI needs besides the map: written once per ROI by roi_tables_kernel (all ROIs of a
// call in parallel, one thread per axis and bin) and copied into shared memory by the pooling CTA.  Without this pre-pass
// the same work is a serial prologue of every CTA - one thread for the box geometry, then 2R threads for the tap lists,
// 5 600 of the 13 900 cycles a CTA lives (tools/prof_roi.py).
template <int R>
struct alignas(16) RoiBlob {
  int xoff[R][kMaxTaps];
  float xw[R][kMaxTaps];
  int yoff[R][kMaxTaps];
  float yw[R][kMaxTaps];
  int xn[R], yn[R];
  float gbox[4];           // start_w, start_h, bin_w, bin_h
  int grid_w, grid_h, W, H;
  int level, tables;
  float inv_count;
  int pad;
};

struct BoxGeom {
  int lvl, H, W, grid_w, grid_h;
  float start_w, start_h, bin_w, bin_h;
};

template <int R>
__device__ __forceinline__ BoxGeom box_geometry(const RoiParams& prm, const float4 box) {
  BoxGeom g;
  const int min_level = 31 - __clz(prm.stride[0]);
  g.lvl = assign_level(box, min_level, prm.num_levels);
  g.H = prm.H[0];
  g.W = prm.W[0];
  int stride = prm.stride[0];
  if (g.lvl == 1) { g.H = prm.H[1]; g.W = prm.W[1]; stride = prm.stride[1]; }
  if (g.lvl == 2) { g.H = prm.H[2]; g.W = prm.W[2]; stride = prm.stride[2]; }
  const float scale = 1.0f / (float)stride;
  g.start_w = __fsub_rn(__fmul_rn(box.x, scale), 0.5f);
  g.start_h = __fsub_rn(__fmul_rn(box.y, scale), 0.5f);
  const float end_w = __fsub_rn(__fmul_rn(box.z, scale), 0.5f);
  const float end_h = __fsub_rn(__fmul_rn(box.w, scale), 0.5f);
  const float roi_w = __fsub_rn(end_w, g.start_w), roi_h = __fsub_rn(end_h, g.start_h);
  g.grid_h = (int)ceilf(__fdiv_rn(roi_h, (float)R));
  g.grid_w = (int)ceilf(__fdiv_rn(roi_w, (float)R));
  g.bin_w = __fdiv_rn(roi_w, (float)R);
  g.bin_h = __fdiv_rn(roi_h, (float)R);
  return g;
}

// One thread per (ROI, axis, bin): grid.x covers roi_cap * 2R threads of a problem, grid.y = problem.
template <int R>
__global__ void __launch_bounds__(256) roi_tables_kernel(RoiParams prm, const float* __restrict__ rois,
                                                         const int32_t* __restrict__ roi_count, RoiBlob<R>* __restrict__ blobs,
                                                         int32_t* __restrict__ out_level) {
  const int p = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t / (2 * R), k = t - r * 2 * R;      // k < R: x bin k, else y bin k - R
  const int cnt = roi_count ? min(roi_count[p], prm.roi_cap) : prm.roi_cap;
  if (r >= cnt) return;
  const float4 box = *reinterpret_cast<const float4*>(rois + ((size_t)p * prm.roi_cap + r) * 4);
  const BoxGeom g = box_geometry<R>(prm, box);
  RoiBlob<R>& B = blobs[(size_t)p * prm.roi_cap + r];
  const bool tables = g.grid_h <= kMaxGrid && g.grid_w <= kMaxGrid;
  if (k == 0) {
    if (out_level) out_level[(size_t)p * prm.roi_cap + r] = g.lvl;
    B.gbox[0] = g.start_w; B.gbox[1] = g.start_h; B.gbox[2] = g.bin_w; B.gbox[3] = g.bin_h;
    B.grid_w = g.grid_w; B.grid_h = g.grid_h; B.W = g.W; B.H = g.H;
    B.level = g.lvl;
    B.tables = tables ? 1 : 0;
    B.inv_count = 1.0f / fmaxf((float)(g.grid_h * g.grid_w), 1.0f);
    B.pad = 0;
  }
  if (!tables) return;
  const bool is_x = k < R;
  const int bin = is_x ? k : k - R;
  // the list of DISTINCT rows / columns the bin's samples touch with their summed weights (see build_axis_taps)
  int idx[kMaxTaps];
  float w[kMaxTaps];
  int n = 0;
  const int grid = is_x ? g.grid_w : g.grid_h, size = is_x ? g.W : g.H;
  const float start = is_x ? g.start_w : g.start_h, bsz = is_x ? g.bin_w : g.bin_h;
  for (int i = 0; i < grid; ++i) {
    const AxisSample s = axis_sample(start, bsz, bin, i, grid, size);
    if (s.w_lo == 0.f && s.w_hi == 0.f) continue;
    if (n > 0 && idx[n - 1] == s.lo) {
      w[n - 1] += s.w_lo;
    } else if (n > 1 && idx[n - 2] == s.lo) {
      w[n - 2] += s.w_lo;
    } else {
      idx[n] = s.lo;
      w[n] = s.w_lo;
      ++n;
    }
    if (idx[n - 1] == s.hi) {
      w[n - 1] += s.w_hi;
    } else {
      idx[n] = s.hi;
      w[n] = s.w_hi;
      ++n;
    }
  }
  const int pitch = is_x ? prm.pix : g.W * prm.pix;
  int* po = is_x ? B.xoff[bin] : B.yoff[bin];
  float* pw = is_x ? B.xw[bin] : B.yw[bin];
#pragma unroll
  for (int q = 0; q < kMaxTaps; ++q) {
    po[q] = (q < n ? idx[q] : 0) * pitch;
    pw[q] = q < n ? w[q] : 0.f;
  }
  (is_x ? B.xn : B.yn)[bin] = n;
}

// NX = padded number of column taps per bin (compile time): the inner loop is 2 LDS.128 per bin and row plus
// LDG.128 + FMUL + 4 FFMA per tap.
template <int R, int NX>
__device__ __forceinline__ void roi_accumulate(const RoiBlob<R>& B, const float* __restrict__ f, int ph, int pw0, int ny,
                                               float4 (&acc)[R / 2]) {
  for (int kr = 0; kr < ny; ++kr) {
    const float wy = B.yw[ph][kr];
    const float* rowp = f + B.yoff[ph][kr];
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      int off[kMaxTaps];
      float w[kMaxTaps];
      *reinterpret_cast<int4*>(off) = *reinterpret_cast<const int4*>(&B.xoff[pw0 + j][0]);
      *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(&B.xw[pw0 + j][0]);
      if (NX > 4) {
        *reinterpret_cast<int4*>(off + 4) = *reinterpret_cast<const int4*>(&B.xoff[pw0 + j][4]);
        *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(&B.xw[pw0 + j][4]);
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

// One CTA per ROI (and 128-channel block); warp = (bin row, half of the bin columns), lane = 4 channels: every tap is
// one fully coalesced 512-byte row of the NHWC map, 4 accumulators per lane keep the register count low enough for
// full occupancy.  The ROI's geometry and tap lists come from roi_tables_kernel.
template <int R>
__global__ void __launch_bounds__(R * 64, R > 8 ? 1 : 2)
roi_align_kernel(RoiParams prm, const int32_t* __restrict__ roi_count, const RoiBlob<R>* __restrict__ blobs,
                 float* __restrict__ pooled) {
  __shared__ RoiBlob<R> B;
  const int r = blockIdx.x, p = blockIdx.y;
#ifdef FOD_ROI_PROF
  const long long t_start = clock64();
#endif
  const int cnt = roi_count ? min(roi_count[p], prm.roi_cap) : prm.roi_cap;
  if (r >= cnt) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const int4* src = reinterpret_cast<const int4*>(blobs + ((size_t)p * prm.roi_cap + r));
    int4* dst = reinterpret_cast<int4*>(&B);
    for (int i = threadIdx.x; i < (int)(sizeof(RoiBlob<R>) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  RPROF(1);
  const bool tables = B.tables != 0;
  const int W = B.W, H = B.H, grid_w = B.grid_w, grid_h = B.grid_h;
  const float* fbase = prm.feat[0];
  if (B.level == 1) fbase = prm.feat[1];
  if (B.level == 2) fbase = prm.feat[2];
  const float* f = fbase + (size_t)(p / prm.C) * H * W * prm.pix + blockIdx.z * kC + lane * 4;   // this CTA's 128-channel block
  const int ph = warp >> 1, pw0 = (warp & 1) * (R / 2);
  float4 acc[R / 2];
#pragma unroll
  for (int j = 0; j < R / 2; ++j) acc[j] = make_float4(0, 0, 0, 0);
  if (tables) {
    int nx = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) nx = max(nx, B.xn[j]);
    const int ny = B.yn[ph];
    switch (nx) {
      case 0: break;
      case 1: case 2: roi_accumulate<R, 2>(B, f, ph, pw0, ny, acc); break;
      case 3: roi_accumulate<R, 3>(B, f, ph, pw0, ny, acc); break;
      case 4: roi_accumulate<R, 4>(B, f, ph, pw0, ny, acc); break;
      case 5: case 6: roi_accumulate<R, 6>(B, f, ph, pw0, ny, acc); break;
      default: roi_accumulate<R, 8>(B, f, ph, pw0, ny, acc); break;
    }
  } else {  // sampling grid beyond the tables (box much larger than the pyramid level expects): sample by sample
    for (int iy = 0; iy < grid_h; ++iy) {
      const AxisSample ys = axis_sample(B.gbox[1], B.gbox[3], ph, iy, grid_h, H);
      if (ys.w_lo == 0.f && ys.w_hi == 0.f) continue;
      const float* row_lo = f + (size_t)ys.lo * W * prm.pix;
      const float* row_hi = f + (size_t)ys.hi * W * prm.pix;
#pragma unroll
      for (int j = 0; j < R / 2; ++j) {
        for (int ix = 0; ix < grid_w; ++ix) {
          const AxisSample xs = axis_sample(B.gbox[0], B.gbox[2], pw0 + j, ix, grid_w, W);
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
  RPROF(2);
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
  const float inv_count = B.inv_count;
#pragma unroll
  for (int j = 0; j < R / 2; ++j) {
    float4 o = make_float4(acc[j].x * inv_count, acc[j].y * inv_count, acc[j].z * inv_count, acc[j].w * inv_count);
    *reinterpret_cast<float4*>(out + (size_t)(ph * R + pw0 + j) * bin_stride) = o;
  }
  RPROF(3);
}

}  // namespace fod

#ifdef FOD_ROI_PROF
extern "C" int fod_roi_prof(long long* out8, int reset) {
  if (reset) {
    long long z[8] = {0};
    cudaMemcpyToSymbol(fod::g_roi_prof, z, sizeof(z));
  } else {
    cudaMemcpyFromSymbol(out8, fod::g_roi_prof, sizeof(long long) * 8);
  }
  return 0;
}
#endif

using namespace fod;

extern "C" int fod_roi_align_wide(const float* const* feat, const fod_level_t* levels, int num_levels, int batch,
                                  int problems_per_image, const float* rois, const int32_t* roi_count, int roi_cap,
                                  int resolution, int channels, int tiled, float* pooled, int32_t* out_level,
                                  void* workspace, fod_stream_t stream) {
  FOD_REQUIRE(feat && levels && rois && pooled && workspace, "fod_roi_align: null pointer");
  FOD_REQUIRE(((uintptr_t)workspace & 15) == 0, "fod_roi_align: workspace must be 16-byte aligned");
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
  const dim3 tgrid((unsigned)((roi_cap * 2 * resolution + 255) / 256), (unsigned)P);
  cudaStream_t st = as_stream(stream);
  if (resolution == 8) {
    auto* blobs = static_cast<RoiBlob<8>*>(workspace);
    roi_tables_kernel<8><<<tgrid, 256, 0, st>>>(prm, rois, roi_count, blobs, out_level);
    roi_align_kernel<8><<<grid, 512, 0, st>>>(prm, roi_count, blobs, pooled);
  } else if (resolution == 4) {
    auto* blobs = static_cast<RoiBlob<4>*>(workspace);
    roi_tables_kernel<4><<<tgrid, 256, 0, st>>>(prm, rois, roi_count, blobs, out_level);
    roi_align_kernel<4><<<grid, 256, 0, st>>>(prm, roi_count, blobs, pooled);
  } else {
    auto* blobs = static_cast<RoiBlob<14>*>(workspace);
    roi_tables_kernel<14><<<tgrid, 256, 0, st>>>(prm, rois, roi_count, blobs, out_level);
    roi_align_kernel<14><<<grid, 896, 0, st>>>(prm, roi_count, blobs, pooled);
  }
  FOD_CUDA_LAUNCH_CHECK("fod_roi_align");
  return FOD_OK;
}

extern "C" size_t fod_roi_align_workspace_bytes(int num_problems, int roi_cap, int resolution) {
  const size_t per = resolution == 8 ? sizeof(RoiBlob<8>) : (resolution == 4 ? sizeof(RoiBlob<4>) : sizeof(RoiBlob<14>));
  return (size_t)num_problems * roi_cap * per;
}

extern "C" int fod_roi_align(const float* const* feat, const fod_level_t* levels, int num_levels, int batch,
                             int problems_per_image, const float* rois, const int32_t* roi_count, int roi_cap,
                             int resolution, int tiled, float* pooled, int32_t* out_level, void* workspace,
                             fod_stream_t stream) {
  return fod_roi_align_wide(feat, levels, num_levels, batch, problems_per_image, rois, roi_count, roi_cap, resolution, kC,
                            tiled, pooled, out_level, workspace, stream);
}
