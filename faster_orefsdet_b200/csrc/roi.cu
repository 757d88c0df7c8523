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

__global__ void __launch_bounds__(256)
roi_align_kernel(RoiParams prm, const float* __restrict__ rois, const int32_t* __restrict__ roi_count,
                 float* __restrict__ pooled, int32_t* __restrict__ out_level) {
  const int r = blockIdx.x, p = blockIdx.y;
  const int cnt = roi_count ? min(roi_count[p], prm.roi_cap) : prm.roi_cap;
  if (r >= cnt) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float4 box = *reinterpret_cast<const float4*>(rois + ((size_t)p * prm.roi_cap + r) * 4);
  int min_level = 31 - __clz(prm.stride[0]);
  const int lvl = assign_level(box, min_level, prm.num_levels);
  if (out_level && threadIdx.x == 0) out_level[(size_t)p * prm.roi_cap + r] = lvl;
  const int H = prm.H[lvl], W = prm.W[lvl];
  const float scale = 1.0f / (float)prm.stride[lvl];
  const float* f = prm.feat[lvl] + (size_t)(p / prm.C) * H * W * kC + lane * 4;
  const int R = prm.R;
  const float start_w = __fsub_rn(__fmul_rn(box.x, scale), 0.5f);
  const float start_h = __fsub_rn(__fmul_rn(box.y, scale), 0.5f);
  const float end_w = __fsub_rn(__fmul_rn(box.z, scale), 0.5f);
  const float end_h = __fsub_rn(__fmul_rn(box.w, scale), 0.5f);
  const float roi_w = __fsub_rn(end_w, start_w), roi_h = __fsub_rn(end_h, start_h);
  const float bin_h = __fdiv_rn(roi_h, (float)R), bin_w = __fdiv_rn(roi_w, (float)R);
  const int grid_h = (int)ceilf(__fdiv_rn(roi_h, (float)R));
  const int grid_w = (int)ceilf(__fdiv_rn(roi_w, (float)R));
  const float count = fmaxf((float)(grid_h * grid_w), 1.0f);
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
    out = pooled + ((size_t)p * prm.roi_cap + r) * R * R * kC + lane * 4;
    bin_stride = kC;
  }
  for (int bin = warp; bin < R * R; bin += nwarps) {
    const int ph = bin / R, pw = bin - ph * R;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int iy = 0; iy < grid_h; ++iy) {
      // y = start_h + ph*bin_h + (iy+.5)*bin_h/grid_h     (left-to-right like the C++ expression)
      float yy = __fadd_rn(__fadd_rn(start_h, __fmul_rn((float)ph, bin_h)),
                           __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), bin_h), (float)grid_h));
      for (int ix = 0; ix < grid_w; ++ix) {
        float xx = __fadd_rn(__fadd_rn(start_w, __fmul_rn((float)pw, bin_w)),
                             __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), bin_w), (float)grid_w));
        float y = yy, x = xx;
        if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
        if (y <= 0.f) y = 0.f;
        if (x <= 0.f) x = 0.f;
        int y_low = (int)y, x_low = (int)x, y_high, x_high;
        if (y_low >= H - 1) {
          y_high = y_low = H - 1;
          y = (float)y_low;
        } else {
          y_high = y_low + 1;
        }
        if (x_low >= W - 1) {
          x_high = x_low = W - 1;
          x = (float)x_low;
        } else {
          x_high = x_low + 1;
        }
        const float ly = __fsub_rn(y, (float)y_low), lx = __fsub_rn(x, (float)x_low);
        const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
        const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
        const float4 v1 = ldg4(f + ((size_t)y_low * W + x_low) * kC);
        const float4 v2 = ldg4(f + ((size_t)y_low * W + x_high) * kC);
        const float4 v3 = ldg4(f + ((size_t)y_high * W + x_low) * kC);
        const float4 v4 = ldg4(f + ((size_t)y_high * W + x_high) * kC);
        acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
        acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
        acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
        acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
      }
    }
    acc.x /= count; acc.y /= count; acc.z /= count; acc.w /= count;
    *reinterpret_cast<float4*>(out + (size_t)bin * bin_stride) = acc;
  }
}

}  // namespace fod

using namespace fod;

extern "C" int fod_roi_align(const float* const* feat, const fod_level_t* levels, int num_levels, int batch,
                             int problems_per_image, const float* rois, const int32_t* roi_count, int roi_cap,
                             int resolution, int tiled, float* pooled, int32_t* out_level, fod_stream_t stream) {
  FOD_REQUIRE(feat && levels && rois && pooled, "fod_roi_align: null pointer");
  FOD_REQUIRE(num_levels >= 1 && num_levels <= FOD_MAX_LEVELS, "fod_roi_align: num_levels %d out of range", num_levels);
  FOD_REQUIRE(batch >= 0 && problems_per_image > 0 && roi_cap > 0 && resolution > 0 && resolution <= 16,
              "fod_roi_align: bad sizes");
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
  FOD_REQUIRE(!tiled || resolution == 8, "fod_roi_align: the tiled layout needs resolution 8");
  prm.tiled = tiled ? 1 : 0;
  dim3 grid(roi_cap, (unsigned)P);
  roi_align_kernel<<<grid, 256, 0, as_stream(stream)>>>(prm, rois, roi_count, pooled, out_level);
  FOD_CUDA_LAUNCH_CHECK("fod_roi_align");
  return FOD_OK;
}

