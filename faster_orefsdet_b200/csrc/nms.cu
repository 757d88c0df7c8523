// Greedy NMS kernels: one problem per CTA or per thread-block CLUSTER, everything resident in shared memory.
//
//   sort   : bitonic sort of 64-bit keys (~ordered(score) << 32 | index) -> stable
//            score-descending order (lower input index first among equal scores).
//   sweep  : blocked greedy suppression, 64 sorted boxes per step:
//              1. 64x64 IoU bit matrix of the chunk, stored by COLUMN (warp ballots),
//              2. the chunk is resolved by ONE WARP with a fixed-point iteration on the column masks
//                 (kept_j = alive_j and no kept_i, i < j, overlaps j: converges to the greedy answer in
//                 (longest suppression chain + 1) ballot rounds instead of one serial step per survivor),
//              3. the chunk's survivors suppress every later box (all threads; four lanes share a box).
//            Only survivors ever test later boxes, so the work is kept*n/2 IoU
//            pairs instead of n^2/2, and no n x n mask goes through HBM.
//            The sweep stops as soon as the caller's top-k is decided.
//   cluster: the kernel is instruction bound on one SM (profiles/r1_ncu_v5_summary.md: 59 % issue-active on 64 of
//            148 SMs), so a problem may be spread over S = 1..8 CTAs of a cluster.  Every CTA holds the sorted boxes
//            and resolves every chunk redundantly (identical results by construction); step 3 is split: CTA r owns
//            the suppression words w with w % S == r.  After each chunk the owners of the next chunk's two words push
//            them into every peer's shared memory (st.shared::cluster) and the cluster barriers once.
//
// IoU arithmetic is bit-identical to torchvision's CPU kernel (common.cuh).
#include "common.cuh"
#include "tc05.cuh"

namespace fod {

#ifdef FOD_NMS_PROF
__device__ long long g_nms_prof[16];
#define PROF_DECL long long prof_acc[16] = {0}
#define PROF_T(var) long long var = clock64()
#define PROF_ADD(slot, t0) prof_acc[slot] += clock64() - (t0)
#define PROF_CNT(slot, v) prof_acc[slot] += (v)
#define PROF_FLUSH(lo, hi) do { if (threadIdx.x == 0 && blockIdx.x == 0) for (int i_ = lo; i_ < hi; ++i_) g_nms_prof[i_] += prof_acc[i_]; } while (0)
#else
#define PROF_DECL
#define PROF_T(var)
#define PROF_ADD(slot, t0)
#define PROF_CNT(slot, v)
#define PROF_FLUSH(lo, hi)
#endif

constexpr int kNmsThreads = 1024;
constexpr int kChunk = 64;
constexpr int kPreRead = 3;      // list passes a thread may hold in registers (suppression pass of greedy_sweep)
constexpr int kMaxCluster = 4;   // measured: 8 CTAs per problem are slower than 4 (cluster barrier + the redundant sort / resolve)

struct NmsSmem {
  unsigned long long* keys;     // [npad]
  float4* box;                  // [ncap] sorted boxes (possibly class-offset)
  float4* kbox;                 // [128] survivors of the chunk pair being applied
  uint32_t* colmask;            // [2 buffers][lo 64 | hi 64]: bit i of word j: chunk box i (i < j) overlaps chunk box j
  float* karea;                 // [128]
  uint32_t* suppressed;         // [ncap/32 + 2]; a CTA of a cluster keeps only its own words up to date
  uint32_t* chunk_sup;          // [2 parities][4 + 2]: suppression words of the chunk pair being resolved (pushed by their
                                // owners), then the pair's two cross words (second chunk suppressed by the first one's survivors)
  int* scalars;                 // [8]
  uint16_t* kept;               // [ncap] sorted ranks of survivors, growing from the front ...
  uint16_t* alist_top;          // ... and, from the END of the same array downwards (entry e at alist_top[-e]), this CTA's
                                // boxes that are still alive and not yet resolved (any order).  While a pair of chunks is
                                // being resolved the list still holds the pair's own <= 128 boxes: survivors + listed
                                // <= n + 128, which is the size of the array.
};

__host__ __device__ inline int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

__host__ __device__ inline size_t nms_words(int ncap) { return (size_t)(ncap + 31) / 32 + 4; }   // 4 spare (zero) words

__host__ inline size_t nms_smem_bytes(int ncap) {
  int npad = next_pow2(ncap < 64 ? 64 : ncap);
  return (size_t)npad * 8 + (size_t)ncap * 16 + 128 * 16 + 2 * 128 * 4 + 128 * 4 + nms_words(ncap) * 4 + 48 + 32 + (size_t)(ncap + 2 * kChunk) * 2 + 64;
}

__device__ inline NmsSmem carve(unsigned char* base, int ncap) {
  int npad = next_pow2(ncap < 64 ? 64 : ncap);
  NmsSmem s;
  s.keys = reinterpret_cast<unsigned long long*>(base);
  base += (size_t)npad * 8;
  s.box = reinterpret_cast<float4*>(base);
  base += (size_t)ncap * 16;
  s.kbox = reinterpret_cast<float4*>(base);
  base += 128 * 16;
  s.colmask = reinterpret_cast<uint32_t*>(base);
  base += 2 * 128 * 4;
  s.karea = reinterpret_cast<float*>(base);
  base += 128 * 4;
  s.suppressed = reinterpret_cast<uint32_t*>(base);
  base += nms_words(ncap) * 4;
  s.chunk_sup = reinterpret_cast<uint32_t*>(base);
  base += 48;
  s.scalars = reinterpret_cast<int*>(base);
  base += 32;
  s.kept = reinterpret_cast<uint16_t*>(base);
  s.alist_top = s.kept + (ncap + 2 * kChunk - 1);
  return s;
}

// One compare-exchange step of the bitonic network on E consecutive keys per thread held in registers (thread t owns
// the global indices E*t .. E*t+E-1): partner distance j < 32*E, so the partner sits in this thread or in another lane
// of the warp (two 32-bit shuffles per key).
template <int E>
__device__ __forceinline__ void bitonic_reg_step(unsigned long long (&v)[E], int k, int j, int g0) {
  if (j < E) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if ((e & j) == 0 && (e | j) < E) {
        const bool up = ((g0 + e) & k) == 0;
        const unsigned long long a = v[e], b = v[e | j];
        if ((a > b) == up) {
          v[e] = b;
          v[e | j] = a;
        }
      }
    }
  } else {
    const int lj = j / E;   // lane distance
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const unsigned long long a = v[e];
      const uint32_t plo = __shfl_xor_sync(0xffffffffu, (uint32_t)a, lj);
      const uint32_t phi = __shfl_xor_sync(0xffffffffu, (uint32_t)(a >> 32), lj);
      const unsigned long long b = ((unsigned long long)phi << 32) | plo;
      const int g = g0 + e;
      const bool keep_min = ((g & j) == 0) == ((g & k) == 0);
      v[e] = keep_min ? (a < b ? a : b) : (a > b ? a : b);
    }
  }
}

template <int E>
__device__ inline void bitonic_sort_reg(unsigned long long* keys, int npad) {
  // npad == E * blockDim.x.  Stages with partner distance >= 32*E go through shared memory, the rest stay in registers.
  const int t = threadIdx.x, g0 = E * t;
  unsigned long long v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = keys[g0 + e];
  const int warp_span = 32 * E;
  // k <= warp_span: entirely inside a warp
  for (int k = 2; k <= warp_span; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) bitonic_reg_step<E>(v, k, j, g0);
  for (int k = warp_span * 2; k <= npad; k <<= 1) {
#pragma unroll
    for (int e = 0; e < E; ++e) keys[g0 + e] = v[e];
    __syncthreads();
    for (int j = k >> 1; j >= warp_span; j >>= 1) {
      for (int tt = t; tt < (npad >> 1); tt += blockDim.x) {
        const int i = 2 * tt - (tt & (j - 1)), l = i + j;
        const unsigned long long a = keys[i], b = keys[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) {
          keys[i] = b;
          keys[l] = a;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = keys[g0 + e];
    for (int j = warp_span >> 1; j > 0; j >>= 1) bitonic_reg_step<E>(v, k, j, g0);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) keys[g0 + e] = v[e];
  __syncthreads();
}

// keys[0..npad) ascending.  Called by the whole CTA (blockDim.x == kNmsThreads).
__device__ inline void bitonic_sort(unsigned long long* keys, int npad) {
  if (npad == 2 * kNmsThreads) return bitonic_sort_reg<2>(keys, npad);
  if (npad == 4 * kNmsThreads) return bitonic_sort_reg<4>(keys, npad);
  if (npad == 8 * kNmsThreads) return bitonic_sort_reg<8>(keys, npad);
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

// barrier over the CTA (S == 1) or over the whole cluster, with release / acquire of shared-memory writes
__device__ __forceinline__ void sync_group(int S) {
  if (S > 1) tc::cluster_sync();
  else __syncthreads();
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

// Greedy sweep over s.box[0..n) (sorted).  score_of_rank(r) gives the score of sorted rank r.
// Stops early once `limit` survivors exist and (if keep_ties) no later box can tie the limit-th score.
// S = CTAs per problem (cluster size), rank = this CTA's rank in the cluster: every CTA calls with identical data and
// gets the same survivor list in s.kept; the return value is its length.
//
// Per pair of chunks: [resolve the first: one warp] sync [bit matrix of the second + its boxes against the first one's
// survivors: all warps] sync [resolve the second: one warp] sync [bit matrix of the next pair's first chunk + suppression
// of later boxes by the pair's survivors: all warps] sync [exchange of the next pair's suppression words: cluster only].  A warp that runs alone is latency bound (~5 cycles
// per dependent instruction), so the resolve step works on 32-bit halves and everything that does not depend on it is
// moved into the all-warps phase.
template <typename ScoreFn>
__device__ inline int greedy_sweep(const NmsSmem& s, int n, float thr_f, int limit, bool keep_ties, ScoreFn score_of_rank,
                                   const int S, const int rank) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwords = (n + 31) >> 5;
  const int nchunks = (n + kChunk - 1) / kChunk;
  PROF_DECL;
  for (int w = tid; w < nwords + 4; w += blockDim.x) s.suppressed[w] = 0;
  if (tid < 4) s.chunk_sup[tid] = 0;  // pair 0 (parity 0) starts with nothing suppressed; parity 1 is always pushed
  // this CTA's boxes: ranks j whose word (j >> 5) is congruent to `rank` modulo S; all alive at the start
  for (int q = tid;; q += blockDim.x) {
    const int j = (rank + (q >> 5) * S) * 32 + (q & 31);
    if (j >= n) break;      // (only the last word of the list is partial, so the valid q form a prefix)
    s.alist_top[-q] = (uint16_t)j;
  }
  if (tid == 0) {
    const int full_words = n >> 5, tail = n & 31;      // words 0..full_words-1 are complete
    int my_n = (full_words > rank ? (full_words - rank + S - 1) / S : 0) * 32;
    if (tail && full_words % S == rank) my_n += tail;
    s.scalars[0] = 0;  // kept count
    s.scalars[1] = 0;
    s.scalars[3] = my_n;
    s.scalars[4] = 0;
  }
  // IoU bit matrix of chunk c2 by column into buffer (c2 & 1): warp w handles columns 2w and 2w+1, lanes are the rows
  // i < j.  Columns already known to be suppressed are skipped (their mask is never looked at).
  auto build_colmask = [&](int c2) {
    const int base2 = c2 * kChunk, m2 = min(kChunk, n - base2);
    uint32_t* clo = s.colmask + (c2 & 1) * 128;
    uint32_t* chi = clo + 64;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int j = warp * 2 + rr;
      unsigned lo = 0, hi = 0;
      if (j < m2 && !((s.suppressed[(base2 + j) >> 5] >> ((base2 + j) & 31)) & 1u)) {
        const float4 bj = s.box[base2 + j];
        const float aj = box_area(bj);
        bool p0 = false, p1 = false;
        if (lane < j) {
          const float4 bi = s.box[base2 + lane];
          p0 = iou_exceeds(bi, box_area(bi), bj, aj, thr_f);
        }
        lo = __ballot_sync(0xffffffffu, p0);
        if (j > 32) {
          if (lane + 32 < j) {
            const float4 bi = s.box[base2 + lane + 32];
            p1 = iou_exceeds(bi, box_area(bi), bj, aj, thr_f);
          }
          hi = __ballot_sync(0xffffffffu, p1);
        }
      }
      if (lane == 0) {
        clo[j] = lo;
        chi[j] = hi;
      }
    }
  };
  __syncthreads();
  build_colmask(0);
  sync_group(S);  // also: every CTA of the cluster is running before anyone writes into a peer's shared memory
  // 2. resolution of chunk c by one warp: fixed point of kept_j = alive_j && !(colmask_j & kept), 32-bit halves.  The
  //    chunk's survivors are appended to s.kept and to the pair's survivor boxes from slot `kofs` on.  (warp 0 only)
  auto resolve = [&](int c, uint32_t suplo, uint32_t suphi, int kofs) {
    const int base = c * kChunk;
    const int m = min(kChunk, n - base);
    const uint32_t validlo = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
    const uint32_t validhi = m == 64 ? 0xffffffffu : (m > 32 ? ((1u << (m - 32)) - 1u) : 0u);
    const uint32_t alo = validlo & ~suplo, ahi = validhi & ~suphi;
    const uint32_t* clo = s.colmask + (c & 1) * 128;
    const uint32_t c0 = clo[lane], c1l = clo[lane + 32], c1h = clo[64 + lane + 32];
    const bool a0 = (alo >> lane) & 1u, a1 = (ahi >> lane) & 1u;
    uint32_t klo = alo, khi = ahi;
    for (int it = 0; it < kChunk; ++it) {
      const uint32_t nlo = __ballot_sync(0xffffffffu, a0 && !(c0 & klo));
      const uint32_t nhi = __ballot_sync(0xffffffffu, a1 && !((c1l & klo) | (c1h & khi)));
      PROF_CNT(10, 1);
      if (nlo == klo && nhi == khi) break;
      klo = nlo;
      khi = nhi;
    }
    int cnt = s.scalars[0];
    const uint32_t lt = (1u << lane) - 1u;
    const int nklo = __popc(klo);
    if ((klo >> lane) & 1u) {
      const int pos = __popc(klo & lt);
      const float4 b = s.box[base + lane];
      s.kept[cnt + pos] = (uint16_t)(base + lane);
      s.kbox[kofs + pos] = b;
      s.karea[kofs + pos] = box_area(b);
    }
    if ((khi >> lane) & 1u) {
      const int pos = nklo + __popc(khi & lt);
      const float4 b = s.box[base + lane + 32];
      s.kept[cnt + pos] = (uint16_t)(base + lane + 32);
      s.kbox[kofs + pos] = b;
      s.karea[kofs + pos] = box_area(b);
    }
    const int nk = nklo + __popc(khi);
    cnt += nk;
    __syncwarp();
    if (lane == 0) {
      s.scalars[0] = cnt;
      s.scalars[2] = kofs + nk;   // survivors of the pair so far
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
  };
  // Chunks are taken in PAIRS: the second chunk of a pair gets the first one's suppression from a direct 64 x nk test
  // (two boxes per warp, redundantly in every CTA of a cluster), so the expensive pass over all later boxes - whose cost
  // is mostly its fixed per-warp skeleton, see profiles/r2_ncu_summary.md - runs once per 128 candidates.
  for (int c = 0; c < nchunks; c += 2) {
    const int base = c * kChunk;
    const int par = (c >> 1) & 1;
    uint32_t* csup = s.chunk_sup + par * 6;   // [0..1] chunk c, [2..3] chunk c + 1, [4..5] cross words of chunk c + 1
    PROF_T(t2);
    if (warp == 0) {
      if (lane < 2) csup[4 + lane] = 0;
      resolve(c, S > 1 ? csup[0] : s.suppressed[base >> 5], S > 1 ? csup[1] : s.suppressed[(base >> 5) + 1], 0);
      if (lane == 0) s.scalars[4] = 0;   // survivor counter of the pair's compaction pass
    }
    __syncthreads();
    if (s.scalars[1] || c + 1 == nchunks) break;  // same decision in every CTA of the cluster (identical data)
    {
      // 1'. bit matrix of the pair's second chunk, and its boxes against the first chunk's survivors
      build_colmask(c + 1);
      const int nk0 = s.scalars[2];
      const int base1 = base + kChunk, m1 = min(kChunk, n - base1);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int j = warp * 2 + rr;
        bool hit = false;
        if (j < m1) {
          const float4 bj = s.box[base1 + j];
          const float aj = box_area(bj);
          for (int k = lane; k < nk0; k += 32) hit = hit || iou_exceeds(s.kbox[k], s.karea[k], bj, aj, thr_f);
        }
        if (__any_sync(0xffffffffu, hit) && lane == 0) atomicOr(&csup[4 + (j >> 5)], 1u << (j & 31));
      }
    }
    __syncthreads();
    if (warp == 0) {
      const int w1 = (base + kChunk) >> 5;
      resolve(c + 1, (S > 1 ? csup[2] : s.suppressed[w1]) | csup[4], (S > 1 ? csup[3] : s.suppressed[w1 + 1]) | csup[5], s.scalars[2]);
    }
    __syncthreads();
    PROF_ADD(2, t2);
    PROF_T(t3);
    if (s.scalars[1] || c + 2 >= nchunks) break;
    const int nk = s.scalars[2];
#ifdef FOD_NMS_DBG
    if (tid == 0 && blockIdx.x == 0) printf("pair c=%d base=%d nk=%d nal=%d kept=%d cross=%08x %08x\n", c, base, nk, s.scalars[3], s.scalars[0], csup[4], csup[5]);
#endif
    // 1'. the next pair's first bit matrix (independent of the suppression state, so it shares this all-warps phase)
    build_colmask(c + 2);
    PROF_ADD(14, t3);
    // 3. survivors of this pair suppress later boxes.  This CTA owns the words w with w % S == rank and keeps a list
    //    of its boxes that are alive and beyond the resolved chunks; one pass tests every listed box against the
    //    pair's survivors (G lanes share a box and split the survivor list) and compacts the list in place, so the
    //    work follows the number of boxes still alive, not the number of candidates.
    if (nk > 0) {   // (a pair without survivors changes nothing; its own entries are dropped by the next pass)
      const int lim = base + 2 * kChunk;
      const int nal = s.scalars[3];
      int lg = 0;
      while (lg < 5 && (nal << (lg + 1)) <= kNmsThreads) ++lg;
      const int G = 1 << lg, per_pass = kNmsThreads >> lg;
      const int sub = tid & (G - 1), grp = tid >> lg;
      const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
      // suppression test of list entry j against the pair's survivors (the G lanes of a group split them)
      auto test = [&](int j, bool live) {
        bool hit = false;
        if (live) {
          const float4 bj = s.box[j];
          const float aj = box_area(bj);
          for (int k = sub; k < nk; k += G) {
            if (iou_exceeds(s.kbox[k], s.karea[k], bj, aj, thr_f)) {
              hit = true;
              break;
            }
          }
        }
        return (__ballot_sync(0xffffffffu, hit) & gmask) != 0u;
      };
      if (nal <= kPreRead * per_pass) {
        // The usual case: every thread fetches its entries of all passes first, so after ONE barrier the compaction may
        // write anywhere in the list; a warp then runs its passes back to back and claims its output slots with one
        // atomic (the order of the list is irrelevant: it is a set).
        int js[kPreRead];
#pragma unroll
        for (int q = 0; q < kPreRead; ++q) {
          const int e = q * per_pass + grp;
          js[q] = e < nal ? (int)s.alist_top[-e] : -1;
        }
        __syncthreads();
        unsigned survbits = 0;   // bit q: entry q of this thread survives (set in the group's first lane only)
        int wtot = 0;
        const int wfirst = (warp * 32) >> lg;   // this warp's first entry of pass 0: a warp without entries skips the pass
#pragma unroll
        for (int q = 0; q < kPreRead; ++q) {
          if (q * per_pass + wfirst < nal) {   // uniform over the warp
            const int j = js[q];
            const bool live = j >= lim;
            const bool dead = test(j, live);
            if (live && dead && sub == 0) atomicOr(&s.suppressed[j >> 5], 1u << (j & 31));
            const bool surv = live && !dead && sub == 0;
            survbits |= (surv ? 1u : 0u) << q;
            wtot += __popc(__ballot_sync(0xffffffffu, surv));
          }
        }
        int wbase = 0;
        if (lane == 0 && wtot) wbase = atomicAdd(&s.scalars[4], wtot);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
#pragma unroll
        for (int q = 0; q < kPreRead; ++q) {
          if (q * per_pass + wfirst < nal) {
            const bool surv = (survbits >> q) & 1u;
            const unsigned sm = __ballot_sync(0xffffffffu, surv);
            if (surv) s.alist_top[-(wbase + __popc(sm & ((1u << lane) - 1u)))] = (uint16_t)js[q];
            wbase += __popc(sm);
          }
        }
      } else {
        for (int e0 = 0; e0 < nal; e0 += per_pass) {
          const int e = e0 + grp;
          const int j = e < nal ? (int)s.alist_top[-e] : -1;
          const bool live = j >= lim;
          const bool dead = test(j, live);
          if (live && dead && sub == 0) atomicOr(&s.suppressed[j >> 5], 1u << (j & 31));
          const bool surv = live && !dead && sub == 0;
          const unsigned sm = __ballot_sync(0xffffffffu, surv);
          int wbase = 0;
          if (lane == 0 && sm) wbase = atomicAdd(&s.scalars[4], __popc(sm));
          wbase = __shfl_sync(0xffffffffu, wbase, 0);
          __syncthreads();  // every entry of this pass has been read; the writes stay below the next pass's entries
          if (surv) s.alist_top[-(wbase + __popc(sm & ((1u << lane) - 1u)))] = (uint16_t)j;
        }
      }
    }
    PROF_ADD(15, t3);
    __syncthreads();
    PROF_ADD(3, t3);
    PROF_T(t4);
    if (tid == 0 && nk > 0) s.scalars[3] = s.scalars[4];
    if (S > 1) {
      // hand the next pair's four suppression words to every CTA (parity-alternating slots: the peers may still be
      // reading this pair's slots)
      const int first_word = (base + 2 * kChunk) >> 5;
      if (tid < 4 * S) {
        const int wsel = tid / S, dst = tid - wsel * S;
        const int w = first_word + wsel;
        if (w % S == rank)   // (zero beyond the list: the array has spare words)
          st_cluster_u32(tc::map_to_cta(tc::smem_u32(&s.chunk_sup[(par ^ 1) * 6 + wsel]), (uint32_t)dst), s.suppressed[w]);
      }
      tc::cluster_sync();
    }
    PROF_ADD(4, t4);
    PROF_ADD(5, t2);
    PROF_CNT(9, 1);
  }
  __syncthreads();
  PROF_FLUSH(0, 6);
  PROF_FLUSH(9, 16);
  return s.scalars[0];
}

// ------------------------------------------------------------------------------------------------
// N0: proposals
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_proposals_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ count,
                     int cand_cap, float thr_f, int post_topk, int roi_cap, int64_t* __restrict__ keep,
                     float* __restrict__ out_boxes, float* __restrict__ out_scores, int32_t* __restrict__ out_count,
                     uint32_t* __restrict__ status, const int S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = blockIdx.x / S, rank = blockIdx.x - p * S;   // S consecutive CTAs = one cluster = one problem
  const int tid = threadIdx.x;
  int n = count ? count[p] : cand_cap;
  n = max(0, min(n, cand_cap));
  PROF_T(tk0);
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
#ifdef FOD_NMS_PROF
  if (threadIdx.x == 0 && blockIdx.x == 0) g_nms_prof[6] += clock64() - tk0;
#endif
  PROF_T(tk1);
  int kept = greedy_sweep(s, n, thr_f, post_topk, true, score_of_rank, S, rank);
#ifdef FOD_NMS_PROF
  if (threadIdx.x == 0 && blockIdx.x == 0) g_nms_prof[7] += clock64() - tk1;
#endif
#ifdef FOD_NMS_PROF
  if (threadIdx.x == 0 && blockIdx.x == 0) g_nms_prof[8] += kept;
#endif
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
  if (rank != 0) return;   // every CTA of the cluster holds the same result; the first one writes it
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
                    int64_t* __restrict__ out_rows, int32_t* __restrict__ out_count, uint32_t* __restrict__ status,
                    const int S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int warp_sums[32];
  __shared__ int sh_total;
  __shared__ float sh_red[32];
  const int b = blockIdx.x / S, rank = blockIdx.x - b * S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
  if (overflow && tid == 0 && rank == 0) atomicOr(status, FOD_STATUS_DET_OVERFLOW);
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
  int kept = greedy_sweep(s, n, thr_f, max_det, false, score_of_rank, S, rank);
  if (max_det >= 0) kept = min(kept, max_det);
  // 4. detector_postprocess on the survivors (d2 postprocessing.py:42-59), order preserved.
  const int oh = out_hw ? out_hw[b * 2] : ih, ow = out_hw ? out_hw[b * 2 + 1] : iw;
  const float sx = (float)((double)ow / (double)iw), sy = (float)((double)oh / (double)ih);
  if (warp == 0 && rank == 0) {
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
                   int n, float thr_f, int64_t* __restrict__ keep, int32_t* __restrict__ keep_count, const int S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float sh_red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = blockIdx.x;   // one problem = one cluster of S CTAs
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
  int kept = greedy_sweep(s, n, thr_f, 0, false, score_of_rank, S, rank);
  if (rank != 0) return;
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

// CTAs per problem: the largest power of two <= kMaxCluster that keeps problems * S within the SM count.
static int pick_cluster(long problems) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int S = 1, smax = kMaxCluster;
  if (const char* e = getenv("FOD_NMS_MAX_CLUSTER")) smax = atoi(e) > 0 ? atoi(e) : smax;   // development knob
  while (S * 2 <= smax && problems * (S * 2) <= sms) S *= 2;
  return S;
}

template <typename K, typename... Args>
static int launch_clustered(const char* name, K kernel, long problems, int S, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(problems * S));
  cfg.blockDim = dim3(kNmsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
    return FOD_ERR_CUDA;
  }
  return FOD_OK;
}

}  // namespace fod

#ifdef FOD_NMS_PROF
extern "C" int fod_nms_prof(long long* out16, int reset) {
  if (reset) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(fod::g_nms_prof, z, sizeof(z));
  } else {
    cudaMemcpyFromSymbol(out16, fod::g_nms_prof, sizeof(long long) * 16);
  }
  return 0;
}
#endif

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
  return launch_clustered("fod_nms_proposals", nms_proposals_kernel, num_problems, pick_cluster(num_problems), smem,
                          as_stream(stream), boxes, scores, count, cand_cap, iou_threshold_as_float(iou_thresh), post_topk,
                          roi_cap, keep, out_boxes, out_scores, out_count, status, pick_cluster(num_problems));
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
  // rows per image decide how far a problem is spread: a few hundred rows are a handful of chunks, not worth a cluster
  const int S = rows > 1024 ? pick_cluster(batch) : 1;
  return launch_clustered("fod_final_detect", final_detect_kernel, batch, S, smem, as_stream(stream), det_boxes, det_scores,
                          roi_count, problems_per_image, roi_cap, ncap, score_thresh, iou_threshold_as_float(iou_thresh),
                          max_det, image_hw, out_hw, out_boxes, out_scores, out_classes, out_rows, out_count, status, S);
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
  const int S = n > 512 ? kMaxCluster : 1;
  return launch_clustered("fod_batched_nms", batched_nms_kernel, 1, S, smem, as_stream(stream), boxes, scores, idxs, n,
                          iou_threshold_as_float(iou_thresh), keep, keep_count, S);
}
