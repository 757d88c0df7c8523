// R1/P1: multi-level ROIAlign (NHWC), R2+R3: folded relation head + scoring + box decode.
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
  float* out = pooled + ((size_t)p * prm.roi_cap + r) * R * R * kC + lane * 4;
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
    *reinterpret_cast<float4*>(out + (size_t)bin * kC) = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Relation head, v1 (fp32 CUDA-core contraction).
//   D[rows][128] = pooled[rows][8192] . w_fold[128][8192]^T ; f = relu(D + bias_cls[c])
//   logits = w_out[0:2] f + b ; deltas = w_out[2:6] f + b ; softmax ; apply_deltas
// tile: 64 ROI rows x 128 columns per CTA, 256 threads, 4x8 register tile, K chunks of 16
// with register double buffering.
// ------------------------------------------------------------------------------------------------
constexpr int kRM = 64, kRN = 128, kRK = 16, kRelThreads = 256;
constexpr int kK = 64 * kC;  // 8192
constexpr float kScaleClamp = 4.135166556742356f;  // log(1000/16), d2 box_regression.py:13

__global__ void __launch_bounds__(kRelThreads, 2)
relation_head_kernel(const float* __restrict__ pooled, const float* __restrict__ w_fold, const float* __restrict__ bias_cls,
                     const float* __restrict__ w_out, const float* __restrict__ b_out, const float* __restrict__ rois,
                     const int32_t* __restrict__ roi_count, int C, int roi_cap, float4 reg_w,
                     float* __restrict__ det_boxes, float* __restrict__ det_scores,
                     float* __restrict__ logits_out, float* __restrict__ deltas_out) {
  __shared__ __align__(16) float As[kRK][kRM + 4];
  __shared__ __align__(16) float Ws[kRK][kRN + 4];
  __shared__ __align__(16) float Wo[6][kRN];
  const int p = blockIdx.y, r0 = blockIdx.x * kRM;
  const int cnt = roi_count ? min(roi_count[p], roi_cap) : roi_cap;
  if (r0 >= cnt) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int i = tid; i < 6 * kRN; i += kRelThreads) (&Wo[0][0])[i] = w_out[i];
  const bool rvalid = (r0 + lr) < cnt;
  const float* arow = pooled + ((size_t)p * roi_cap + r0 + lr) * kK + lk;
  const float* wrow0 = w_fold + (size_t)lr * kK + lk;
  const float* wrow1 = w_fold + (size_t)(lr + 64) * kK + lk;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 an = rvalid ? ldg4(arow) : make_float4(0, 0, 0, 0);
  float4 wn0 = ldg4(wrow0), wn1 = ldg4(wrow1);
  for (int k0 = 0; k0 < kK; k0 += kRK) {
    __syncthreads();
    As[lk + 0][lr] = an.x; As[lk + 1][lr] = an.y; As[lk + 2][lr] = an.z; As[lk + 3][lr] = an.w;
    Ws[lk + 0][lr] = wn0.x; Ws[lk + 1][lr] = wn0.y; Ws[lk + 2][lr] = wn0.z; Ws[lk + 3][lr] = wn0.w;
    Ws[lk + 0][lr + 64] = wn1.x; Ws[lk + 1][lr + 64] = wn1.y; Ws[lk + 2][lr + 64] = wn1.z; Ws[lk + 3][lr + 64] = wn1.w;
    __syncthreads();
    if (k0 + kRK < kK) {
      an = rvalid ? ldg4(arow + k0 + kRK) : make_float4(0, 0, 0, 0);
      wn0 = ldg4(wrow0 + k0 + kRK);
      wn1 = ldg4(wrow1 + k0 + kRK);
    }
#pragma unroll
    for (int k = 0; k < kRK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8 + 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }
  // epilogue
  const int c = p % C;
  const float4 bb0 = ldg4(bias_cls + (size_t)c * kRN + tx * 8), bb1 = ldg4(bias_cls + (size_t)c * kRN + tx * 8 + 4);
  const float bv[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float fj = fmaxf(acc[i][j] + bv[j], 0.f);
#pragma unroll
      for (int o = 0; o < 6; ++o) part[o] = fmaf(fj, Wo[o][tx * 8 + j], part[o]);
    }
#pragma unroll
    for (int o = 0; o < 6; ++o) {
#pragma unroll
      for (int s = 8; s > 0; s >>= 1) part[o] += __shfl_xor_sync(0xffffffffu, part[o], s);
    }
    const int r = r0 + ty * 4 + i;
    if (tx == 0 && r < cnt) {
      const size_t row = (size_t)p * roi_cap + r;
      const float l0 = part[0] + b_out[0], l1 = part[1] + b_out[1];
      const float d0 = part[2] + b_out[2], d1 = part[3] + b_out[3], d2 = part[4] + b_out[4], d3 = part[5] + b_out[5];
      if (logits_out) {
        logits_out[row * 2] = l0;
        logits_out[row * 2 + 1] = l1;
      }
      if (deltas_out) *reinterpret_cast<float4*>(deltas_out + row * 4) = make_float4(d0, d1, d2, d3);
      // softmax over (fg, bg) -> fg probability (custom_fast_rcnn.py:169)
      const float m = fmaxf(l0, l1);
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      det_scores[row] = e0 / (e0 + e1);
      // apply_deltas (box_regression.py:87-115); clipping happens in fod_final_detect, after the
      // reference's isfinite filter (d2 fast_rcnn.py:137-147)
      const float4 bx = *reinterpret_cast<const float4*>(rois + row * 4);
      const float w = bx.z - bx.x, h = bx.w - bx.y;
      const float cx = bx.x + 0.5f * w, cy = bx.y + 0.5f * h;
      const float dx = d0 / reg_w.x, dy = d1 / reg_w.y;
      const float dw = fminf(d2 / reg_w.z, kScaleClamp), dh = fminf(d3 / reg_w.w, kScaleClamp);
      const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
      const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
      const float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
      const float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
      *reinterpret_cast<float4*>(det_boxes + row * 4) = make_float4(x1, y1, x2, y2);
    }
  }
}

}  // namespace fod

using namespace fod;

extern "C" int fod_roi_align(const float* const* feat, const fod_level_t* levels, int num_levels, int batch,
                             int problems_per_image, const float* rois, const int32_t* roi_count, int roi_cap,
                             int resolution, float* pooled, int32_t* out_level, fod_stream_t stream) {
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
  dim3 grid(roi_cap, (unsigned)P);
  roi_align_kernel<<<grid, 256, 0, as_stream(stream)>>>(prm, rois, roi_count, pooled, out_level);
  FOD_CUDA_LAUNCH_CHECK("fod_roi_align");
  return FOD_OK;
}

extern "C" int fod_relation_head(const float* pooled, const float* w_fold, const float* bias_cls, const float* w_out,
                                 const float* b_out, const float* rois, const int32_t* roi_count, int num_problems,
                                 int problems_per_image, int roi_cap, const float* reg_weights, float* det_boxes,
                                 float* det_scores, float* logits, float* deltas, fod_stream_t stream) {
  FOD_REQUIRE(pooled && w_fold && bias_cls && w_out && b_out && rois && reg_weights && det_boxes && det_scores,
              "fod_relation_head: null pointer");
  FOD_REQUIRE(num_problems >= 0 && problems_per_image > 0 && roi_cap > 0, "fod_relation_head: bad sizes");
  FOD_REQUIRE(num_problems % problems_per_image == 0, "fod_relation_head: num_problems not a multiple of classes");
  if (num_problems == 0) return FOD_OK;
  FOD_REQUIRE(num_problems <= 65535, "fod_relation_head: num_problems %d > 65535", num_problems);
  dim3 grid((roi_cap + kRM - 1) / kRM, num_problems);
  float4 rw = make_float4(reg_weights[0], reg_weights[1], reg_weights[2], reg_weights[3]);
  relation_head_kernel<<<grid, kRelThreads, 0, as_stream(stream)>>>(pooled, w_fold, bias_cls, w_out, b_out, rois,
                                                                    roi_count, problems_per_image, roi_cap, rw,
                                                                    det_boxes, det_scores, logits, deltas);
  FOD_CUDA_LAUNCH_CHECK("fod_relation_head");
  return FOD_OK;
}
