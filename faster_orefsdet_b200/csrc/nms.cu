// Greedy NMS kernels: one CTA per problem, everything resident in shared memory.
//
//   sort   : bitonic sort of 64-bit keys (~ordered(score) << 32 | index) -> stable
//            score-descending order (lower input index first among equal scores).
//   sweep  : blocked greedy suppression, 64 sorted boxes per step:
//              1. 64x64 IoU bit matrix of the chunk (warp ballots),
//              2. the chunk is resolved serially by one thread on the bit rows,
//              3. the chunk's survivors suppress every later box (all threads).
//            Only survivors ever test later boxes, so the work is kept*n/2 IoU
//            pairs instead of n^2/2, and no n x n mask goes through HBM.
//            The sweep stops as soon as the caller's top-k is decided.
//
// IoU arithmetic is bit-identical to torchvision's CPU kernel (common.cuh).
#include "common.cuh"

namespace fod {

constexpr int kNmsThreads = 1024;
constexpr int kChunk = 64;

struct NmsSmem {
  unsigned long long* keys;  // [npad]
  float4* box;               // [ncap] sorted boxes (possibly class-offset)
  uint32_t* suppressed;      // [ncap/32]
  uint16_t* kept;            // [ncap] sorted ranks of survivors
  unsigned long long* rowmask;  // [64]
  int* scalars;              // [8]
};

__host__ __device__ inline int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

__host__ inline size_t nms_smem_bytes(int ncap) {
  int npad = next_pow2(ncap < 64 ? 64 : ncap);
  return (size_t)npad * 8 + (size_t)ncap * 16 + ((size_t)(ncap + 31) / 32) * 4 + (size_t)ncap * 2 + 64 * 8 + 64 + 64;
}

__device__ inline NmsSmem carve(unsigned char* base, int ncap) {
  int npad = next_pow2(ncap < 64 ? 64 : ncap);
  NmsSmem s;
  s.keys = reinterpret_cast<unsigned long long*>(base);
  base += (size_t)npad * 8;
  s.box = reinterpret_cast<float4*>(base);
  base += (size_t)ncap * 16;
  s.rowmask = reinterpret_cast<unsigned long long*>(base);
  base += 64 * 8;
  s.suppressed = reinterpret_cast<uint32_t*>(base);
  base += ((size_t)(ncap + 31) / 32) * 4;
  s.scalars = reinterpret_cast<int*>(base);
  base += 32;
  s.kept = reinterpret_cast<uint16_t*>(base);
  return s;
}

// keys[0..npad) ascending.  Called by the whole CTA.
__device__ inline void bitonic_sort(unsigned long long* keys, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
        int i = 2 * t - (t & (j - 1));
        int l = i + j;
        unsigned long long a = keys[i], b = keys[l];
        bool up = (i & k) == 0;
        if ((a > b) == up) {
          keys[i] = b;
          keys[l] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Greedy sweep over s.box[0..n) (sorted).  score_of_rank(r) gives the score of sorted rank r.
// Stops early once `limit` survivors exist and (if keep_ties) no later box can tie the limit-th score.
// Returns the number of survivors recorded in s.kept (all threads get the same value).
template <typename ScoreFn>
__device__ inline int greedy_sweep(const NmsSmem& s, int n, float thr_f, int limit, bool keep_ties, ScoreFn score_of_rank) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int w = tid; w < (n + 31) / 32; w += blockDim.x) s.suppressed[w] = 0;
  if (tid == 0) s.scalars[0] = 0;  // kept count
  __syncthreads();
  int nchunks = (n + kChunk - 1) / kChunk;
  for (int c = 0; c < nchunks; ++c) {
    const int base = c * kChunk;
    const int m = min(kChunk, n - base);
    // 1. intra-chunk bit matrix: warp w handles rows 2w and 2w+1
    for (int rr = 0; rr < 2; ++rr) {
      int i = warp * 2 + rr;
      unsigned lo = 0, hi = 0;
      if (i < m) {
        float4 bi = s.box[base + i];
        float ai = box_area(bi);
        int j0 = lane, j1 = lane + 32;
        bool p0 = false, p1 = false;
        if (j0 < m && j0 > i) {
          float4 bj = s.box[base + j0];
          p0 = iou_exceeds(bi, ai, bj, box_area(bj), thr_f);
        }
        if (j1 < m && j1 > i) {
          float4 bj = s.box[base + j1];
          p1 = iou_exceeds(bi, ai, bj, box_area(bj), thr_f);
        }
        lo = __ballot_sync(0xffffffffu, p0);
        hi = __ballot_sync(0xffffffffu, p1);
      } else {
        __ballot_sync(0xffffffffu, false);
        __ballot_sync(0xffffffffu, false);
      }
      if (lane == 0 && i < kChunk) s.rowmask[i] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
    // 2. serial resolution of the chunk
    if (tid == 0) {
      unsigned long long alive = (m == 64) ? ~0ull : ((1ull << m) - 1ull);
      unsigned long long sup = (unsigned long long)s.suppressed[base >> 5];
      if (m > 32) sup |= (unsigned long long)s.suppressed[(base >> 5) + 1] << 32;
      alive &= ~sup;
      unsigned long long keptbits = 0;
      while (alive) {
        int i = __ffsll((long long)alive) - 1;
        unsigned long long bit = 1ull << i;
        keptbits |= bit;
        alive &= ~(s.rowmask[i] | bit);
      }
      int cnt = s.scalars[0];
      s.scalars[2] = (int)(keptbits & 0xffffffffu);
      s.scalars[3] = (int)(keptbits >> 32);
      unsigned long long kb = keptbits;
      while (kb) {
        int i = __ffsll((long long)kb) - 1;
        kb &= kb - 1;
        s.kept[cnt++] = (uint16_t)(base + i);
      }
      s.scalars[0] = cnt;
      // early stop: the top-`limit` survivors are decided
      int stop = 0;
      if (limit > 0 && cnt >= limit && base + kChunk < n) {
        if (!keep_ties) {
          stop = 1;
        } else {
          float sl = score_of_rank(s.kept[limit - 1]);
          if (score_of_rank(base + kChunk) < sl) stop = 1;
        }
      }
      s.scalars[1] = stop;
    }
    __syncthreads();
    if (s.scalars[1]) break;
    unsigned long long keptbits = ((unsigned long long)(unsigned)s.scalars[3] << 32) | (unsigned)s.scalars[2];
    // 3. survivors of this chunk suppress later boxes
    if (keptbits) {
      for (int j = base + kChunk + tid; j < n; j += blockDim.x) {
        if ((s.suppressed[j >> 5] >> (j & 31)) & 1u) continue;
        float4 bj = s.box[j];
        float aj = box_area(bj);
        unsigned long long kb = keptbits;
        bool dead = false;
        while (kb) {
          int i = __ffsll((long long)kb) - 1;
          kb &= kb - 1;
          float4 bi = s.box[base + i];
          if (iou_exceeds(bi, box_area(bi), bj, aj, thr_f)) {
            dead = true;
            break;
          }
        }
        if (dead) atomicOr(&s.suppressed[j >> 5], 1u << (j & 31));
      }
    }
    __syncthreads();
  }
  __syncthreads();
  return s.scalars[0];
}

// ------------------------------------------------------------------------------------------------
// N0: proposals
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_proposals_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ count,
                     int cand_cap, float thr_f, int post_topk, int roi_cap, int64_t* __restrict__ keep,
                     float* __restrict__ out_boxes, float* __restrict__ out_scores, int32_t* __restrict__ out_count,
                     uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = blockIdx.x;
  const int tid = threadIdx.x;
  int n = count ? count[p] : cand_cap;
  n = max(0, min(n, cand_cap));
  NmsSmem s = carve(smem_raw, cand_cap);
  const float* pb = boxes + (size_t)p * cand_cap * 4;
  const float* ps = scores + (size_t)p * cand_cap;
  const int npad = next_pow2(max(n, 64));
  for (int i = tid; i < npad; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n) k = ((unsigned long long)(~float_to_ordered(ps[i])) << 32) | (unsigned)i;
    s.keys[i] = k;
  }
  __syncthreads();
  bitonic_sort(s.keys, npad);
  for (int r = tid; r < n; r += blockDim.x) {
    int idx = (int)(s.keys[r] & 0xffffffffu);
    s.box[r] = *reinterpret_cast<const float4*>(pb + (size_t)idx * 4);
  }
  __syncthreads();
  auto score_of_rank = [&](int r) { return ps[(int)(s.keys[r] & 0xffffffffu)]; };
  int kept = greedy_sweep(s, n, thr_f, post_topk, true, score_of_rank);
  int m = kept;
  if (post_topk > 0 && kept > post_topk) {
    // keep every survivor whose score >= the post_topk-th best (fsod_rpn.py:1198-1206)
    float sl = score_of_rank(s.kept[post_topk - 1]);
    if (tid == 0) s.scalars[4] = kept;
    __syncthreads();
    for (int i = post_topk + tid; i < kept; i += blockDim.x) {
      // first position whose score drops below sl
      bool below = score_of_rank(s.kept[i]) < sl;
      bool prev_below = score_of_rank(s.kept[i - 1]) < sl;
      if (below && !prev_below) s.scalars[4] = i;
    }
    __syncthreads();
    m = s.scalars[4];
  }
  if (m > roi_cap) {
    if (tid == 0) atomicOr(status, FOD_STATUS_PROPOSAL_OVERFLOW);
    m = roi_cap;
  }
  for (int i = tid; i < m; i += blockDim.x) {
    int r = s.kept[i];
    int idx = (int)(s.keys[r] & 0xffffffffu);
    keep[(size_t)p * roi_cap + i] = idx;
    *reinterpret_cast<float4*>(out_boxes + ((size_t)p * roi_cap + i) * 4) = s.box[r];
    out_scores[(size_t)p * roi_cap + i] = ps[idx];
  }
  if (tid == 0) out_count[p] = m;
}

// ------------------------------------------------------------------------------------------------
// R4 / N1 / O1: final class-wise NMS + postprocess, one CTA per image
// ------------------------------------------------------------------------------------------------
// Exclusive scan of one int per thread over the CTA (blockDim.x == 1024).  *total gets the sum.
__device__ inline int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
    int xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    warp_sums[lane] = xs - w;
    if (lane == 31) *total = xs;
  }
  __syncthreads();
  int res = warp_sums[warp] + x - v;
  __syncthreads();  // warp_sums may be rewritten by the next call; *total stays valid until then
  return res;
}

// Shared tail of fod_final_detect / fod_batched_nms.
//   row ids in the low key bits address `boxes4` / `scores` / `cls_of_row`.
struct RowSource {
  const float* boxes;     // [rows][4]
  const float* scores;    // [rows]
  const int64_t* idxs;    // per-row class (generic op) or NULL
  int roi_cap;            // class = row / roi_cap when idxs == NULL and roi_cap > 0
};

__device__ __forceinline__ float4 clip_box(float4 b, float w, float h) {
  return make_float4(fminf(fmaxf(b.x, 0.f), w), fminf(fmaxf(b.y, 0.f), h), fminf(fmaxf(b.z, 0.f), w),
                     fminf(fmaxf(b.w, 0.f), h));
}

__device__ inline int class_of(const RowSource& src, int row) {
  if (src.idxs) return (int)src.idxs[row];
  return src.roi_cap > 0 ? row / src.roi_cap : 0;
}

__global__ void __launch_bounds__(kNmsThreads, 1)
final_detect_kernel(const float* __restrict__ det_boxes, const float* __restrict__ det_scores,
                    const int32_t* __restrict__ roi_count, int C, int roi_cap, int ncap, float score_thresh, float thr_f,
                    int max_det, const int32_t* __restrict__ image_hw, const int32_t* __restrict__ out_hw,
                    float* __restrict__ out_boxes, float* __restrict__ out_scores, int64_t* __restrict__ out_classes,
                    int64_t* __restrict__ out_rows, int32_t* __restrict__ out_count, uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int warp_sums[32];
  __shared__ int sh_total;
  __shared__ float sh_red[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  NmsSmem s = carve(smem_raw, ncap);
  const size_t img_row0 = (size_t)b * C * roi_cap;
  const float* ib = det_boxes + img_row0 * 4;
  const float* is = det_scores + img_row0;
  const int total_rows = C * roi_cap;
  const int ih = image_hw[b * 2], iw = image_hw[b * 2 + 1];
  const float fw = (float)iw, fh = (float)ih;
  // 1. compaction of valid rows in class-major order (d2 fast_rcnn.py:137-155)
  int n = 0;
  float lmax = -INFINITY;
  bool overflow = false;
  for (int start = 0; start < total_rows; start += blockDim.x) {
    int row = start + tid;
    bool ok = false;
    float sc = 0.f;
    if (row < total_rows) {
      int c = row / roi_cap, r = row - c * roi_cap;
      int cnt = roi_count ? min(roi_count[b * C + c], roi_cap) : roi_cap;
      if (r < cnt) {
        sc = is[row];
        float4 bx = *reinterpret_cast<const float4*>(ib + (size_t)row * 4);
        bool fin = isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w) && isfinite(sc);
        ok = fin && (sc > score_thresh);
        if (ok) {
          bx = clip_box(bx, fw, fh);
          lmax = fmaxf(lmax, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
        }
      }
    }
    int pos = block_exclusive_scan(ok ? 1 : 0, warp_sums, &sh_total);
    int tot = sh_total;
    if (ok) {
      int i = n + pos;
      if (i < ncap) s.keys[i] = ((unsigned long long)(~float_to_ordered(sc)) << 32) | (unsigned)row;
    }
    n += tot;
    __syncthreads();
  }
  if (n > ncap) {
    overflow = true;
    n = ncap;
  }
  if (overflow && tid == 0) atomicOr(status, FOD_STATUS_DET_OVERFLOW);
  // 2. max coordinate (torchvision boxes.py batched_nms: boxes.max())
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) sh_red[warp] = lmax;
  __syncthreads();
  if (warp == 0) {
    float v = sh_red[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) sh_red[0] = v;
  }
  __syncthreads();
  const float coord_step = __fadd_rn(sh_red[0], 1.0f);
  // 3. sort, gather class-offset boxes
  const int npad = next_pow2(max(n, 64));
  for (int i = n + tid; i < npad; i += blockDim.x) s.keys[i] = ~0ull;
  __syncthreads();
  bitonic_sort(s.keys, npad);
  for (int r = tid; r < n; r += blockDim.x) {
    int row = (int)(s.keys[r] & 0xffffffffu);
    float4 bx = clip_box(*reinterpret_cast<const float4*>(ib + (size_t)row * 4), fw, fh);
    float off = __fmul_rn((float)(row / roi_cap), coord_step);
    bx.x = __fadd_rn(bx.x, off);
    bx.y = __fadd_rn(bx.y, off);
    bx.z = __fadd_rn(bx.z, off);
    bx.w = __fadd_rn(bx.w, off);
    s.box[r] = bx;
  }
  __syncthreads();
  auto score_of_rank = [&](int r) { return is[(int)(s.keys[r] & 0xffffffffu)]; };
  int kept = greedy_sweep(s, n, thr_f, max_det, false, score_of_rank);
  if (max_det >= 0) kept = min(kept, max_det);
  // 4. detector_postprocess on the survivors (d2 postprocessing.py:42-59), order preserved.
  const int oh = out_hw ? out_hw[b * 2] : ih, ow = out_hw ? out_hw[b * 2 + 1] : iw;
  const float sx = (float)((double)ow / (double)iw), sy = (float)((double)oh / (double)ih);
  if (warp == 0) {
    int written = 0;
    for (int i0 = 0; i0 < kept; i0 += 32) {
      int i = i0 + lane;
      bool ok = false;
      float4 bx = make_float4(0, 0, 0, 0);
      int row = 0;
      if (i < kept) {
        row = (int)(s.keys[s.kept[i]] & 0xffffffffu);
        bx = clip_box(*reinterpret_cast<const float4*>(ib + (size_t)row * 4), fw, fh);
        bx.x = fminf(fmaxf(__fmul_rn(bx.x, sx), 0.f), (float)ow);
        bx.z = fminf(fmaxf(__fmul_rn(bx.z, sx), 0.f), (float)ow);
        bx.y = fminf(fmaxf(__fmul_rn(bx.y, sy), 0.f), (float)oh);
        bx.w = fminf(fmaxf(__fmul_rn(bx.w, sy), 0.f), (float)oh);
        ok = (__fsub_rn(bx.z, bx.x) > 0.f) && (__fsub_rn(bx.w, bx.y) > 0.f);
      }
      unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        int o = written + __popc(m & ((1u << lane) - 1u));
        size_t base = (size_t)b * max_det + o;
        *reinterpret_cast<float4*>(out_boxes + base * 4) = bx;
        out_scores[base] = is[row];
        out_classes[base] = row / roi_cap;
        if (out_rows) out_rows[base] = row;
      }
      written += __popc(m);
    }
    if (lane == 0) out_count[b] = written;
  }
}

// Generic single-problem batched_nms (operator boundary).
__global__ void __launch_bounds__(kNmsThreads, 1)
batched_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ idxs,
                   int n, float thr_f, int64_t* __restrict__ keep, int32_t* __restrict__ keep_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float sh_red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  NmsSmem s = carve(smem_raw, n);
  const int npad = next_pow2(max(n, 64));
  float lmax = -INFINITY;
  for (int i = tid; i < npad; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n) {
      k = ((unsigned long long)(~float_to_ordered(scores[i])) << 32) | (unsigned)i;
      float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)i * 4);
      lmax = fmaxf(lmax, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
    s.keys[i] = k;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) sh_red[warp] = lmax;
  __syncthreads();
  if (warp == 0) {
    float v = sh_red[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) sh_red[0] = v;
  }
  __syncthreads();
  const float coord_step = __fadd_rn(sh_red[0], 1.0f);
  bitonic_sort(s.keys, npad);
  for (int r = tid; r < n; r += blockDim.x) {
    int i = (int)(s.keys[r] & 0xffffffffu);
    float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)i * 4);
    if (idxs) {
      float off = __fmul_rn((float)idxs[i], coord_step);
      bx.x = __fadd_rn(bx.x, off);
      bx.y = __fadd_rn(bx.y, off);
      bx.z = __fadd_rn(bx.z, off);
      bx.w = __fadd_rn(bx.w, off);
    }
    s.box[r] = bx;
  }
  __syncthreads();
  auto score_of_rank = [&](int r) { return scores[(int)(s.keys[r] & 0xffffffffu)]; };
  int kept = greedy_sweep(s, n, thr_f, 0, false, score_of_rank);
  for (int i = tid; i < kept; i += blockDim.x) keep[i] = (int64_t)(s.keys[s.kept[i]] & 0xffffffffu);
  if (tid == 0) *keep_count = kept;
}

template <typename K>
static int set_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 227 * 1024) {
    set_error("%s: needs %zu bytes of shared memory (> 227 KB)", name, bytes);
    return FOD_ERR_CAPACITY;
  }
  FOD_CUDA_CALL(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return FOD_OK;
}

}  // namespace fod

using namespace fod;

extern "C" int fod_nms_proposals(const float* boxes, const float* scores, const int32_t* count, int num_problems,
                                 int cand_cap, double iou_thresh, int post_topk, int roi_cap, int64_t* keep,
                                 float* out_boxes, float* out_scores, int32_t* out_count, uint32_t* status,
                                 fod_stream_t stream) {
  FOD_REQUIRE(boxes && scores && keep && out_boxes && out_scores && out_count && status, "fod_nms_proposals: null pointer");
  FOD_REQUIRE(num_problems >= 0 && cand_cap > 0 && roi_cap > 0, "fod_nms_proposals: bad sizes");
  if (cand_cap > FOD_NMS_MAX_BOXES) {
    set_error("fod_nms_proposals: cand_cap %d > %d", cand_cap, FOD_NMS_MAX_BOXES);
    return FOD_ERR_CAPACITY;
  }
  if (num_problems == 0) return FOD_OK;
  size_t smem = nms_smem_bytes(cand_cap);
  int rc = set_smem(nms_proposals_kernel, smem, "fod_nms_proposals");
  if (rc != FOD_OK) return rc;
  nms_proposals_kernel<<<num_problems, kNmsThreads, smem, as_stream(stream)>>>(
      boxes, scores, count, cand_cap, iou_threshold_as_float(iou_thresh), post_topk, roi_cap, keep, out_boxes,
      out_scores, out_count, status);
  FOD_CUDA_LAUNCH_CHECK("fod_nms_proposals");
  return FOD_OK;
}

extern "C" int fod_final_detect(const float* det_boxes, const float* det_scores, const int32_t* roi_count, int batch,
                                int problems_per_image, int roi_cap, float score_thresh, double iou_thresh, int max_det,
                                const int32_t* image_hw, const int32_t* out_hw, float* out_boxes, float* out_scores,
                                int64_t* out_classes, int64_t* out_rows, int32_t* out_count, uint32_t* status,
                                fod_stream_t stream) {
  FOD_REQUIRE(det_boxes && det_scores && image_hw && out_boxes && out_scores && out_classes && out_count && status,
              "fod_final_detect: null pointer");
  FOD_REQUIRE(batch >= 0 && problems_per_image > 0 && roi_cap > 0 && max_det > 0, "fod_final_detect: bad sizes");
  if (batch == 0) return FOD_OK;
  long rows = (long)problems_per_image * roi_cap;
  int ncap = (int)(rows < FOD_NMS_MAX_BOXES ? rows : FOD_NMS_MAX_BOXES);
  size_t smem = nms_smem_bytes(ncap);
  int rc = set_smem(final_detect_kernel, smem, "fod_final_detect");
  if (rc != FOD_OK) return rc;
  final_detect_kernel<<<batch, kNmsThreads, smem, as_stream(stream)>>>(
      det_boxes, det_scores, roi_count, problems_per_image, roi_cap, ncap, score_thresh,
      iou_threshold_as_float(iou_thresh), max_det, image_hw, out_hw, out_boxes, out_scores, out_classes, out_rows,
      out_count, status);
  FOD_CUDA_LAUNCH_CHECK("fod_final_detect");
  return FOD_OK;
}

extern "C" int fod_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int n, double iou_thresh,
                               int64_t* keep, int32_t* keep_count, fod_stream_t stream) {
  FOD_REQUIRE(keep_count, "fod_batched_nms: null keep_count");
  FOD_REQUIRE(n >= 0, "fod_batched_nms: negative n");
  if (n > FOD_NMS_MAX_BOXES) {
    set_error("fod_batched_nms: n %d > %d", n, FOD_NMS_MAX_BOXES);
    return FOD_ERR_CAPACITY;
  }
  if (n == 0) {
    FOD_CUDA_CALL(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), as_stream(stream)));
    return FOD_OK;
  }
  FOD_REQUIRE(boxes && scores && keep, "fod_batched_nms: null pointer");
  size_t smem = nms_smem_bytes(n);
  int rc = set_smem(batched_nms_kernel, smem, "fod_batched_nms");
  if (rc != FOD_OK) return rc;
  batched_nms_kernel<<<1, kNmsThreads, smem, as_stream(stream)>>>(boxes, scores, idxs, n,
                                                                   iou_threshold_as_float(iou_thresh), keep, keep_count);
  FOD_CUDA_LAUNCH_CHECK("fod_batched_nms");
  return FOD_OK;
}
