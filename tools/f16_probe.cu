// Probe for the fp16-split variant of the tensor-core GEMM: tcgen05.mma kind::f16 with A in tensor memory (packed
// half2 per 32-bit column), B in shared memory (fp16, K-major, 64-byte swizzle, loaded by TMA).  x = hi + lo with
// hi = fp16(x * 2^e), lo = fp16(x * 2^e - hi): three products hi.hi + lo.hi + hi.lo, fp32 accumulation.
// Development tool:  f16_probe [N]
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../faster_orefsdet_b200/csrc/common.cuh"
#include "../faster_orefsdet_b200/csrc/tc05.cuh"

using namespace fod;
using namespace fod::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

constexpr int K = 32;

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t smem_desc_k_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;   // 8 rows x 64 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;            // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct P { CUtensorMap bhi, blo; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ P p, const float* __restrict__ X, float xs, float* __restrict__ out,
                                                int N, int repeat) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sb = smem_u32(smem), plane = (uint32_t)N * 64;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&tb), 512); tmem_relinquish<1>(); }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); fence_barrier_init();
    mbar_arrive_expect_tx(smem_u32(&bar2), 2 * plane);
    tma_load_2d(sb, &p.bhi, smem_u32(&bar2), 0, 0);
    tma_load_2d(sb + plane, &p.blo, smem_u32(&bar2), 0, 0);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t0 = tb;
  // A row tid: packed half2 columns: hi at [0,16), lo at [16,32)
  uint32_t vh[16], vl[16];
  for (int j = 0; j < 16; ++j) {
    float a = X[tid * K + 2 * j] * xs, b = X[tid * K + 2 * j + 1] * xs;
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    vh[j] = *reinterpret_cast<uint32_t*>(&h);
    vl[j] = *reinterpret_cast<uint32_t*>(&l);
  }
  const uint32_t ta = t0 + ((uint32_t)(warp * 32) << 16);
  tmem_st16(ta, vh); tmem_st16(ta + 16, vl); tmem_wait_st();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  mbar_wait(smem_u32(&bar2), 0);
  long long t_start = clock64();
  if (warp == 0) {
    const uint32_t idesc = idesc_f16(128, N), d = t0 + 64;
    const uint64_t bh = smem_desc_k_sw64(sb), bl = smem_desc_k_sw64(sb + plane);
    for (int rep = 0; rep < repeat; ++rep) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t bo = (uint64_t)((ks * 32) >> 4);
          mma_f16_ts(d, t0 + ks * 8, bh + bo, idesc, (rep | ks) ? 1u : 0u);
          mma_f16_ts(d, t0 + 16 + ks * 8, bh + bo, idesc, 1u);
          mma_f16_ts(d, t0 + ks * 8, bl + bo, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (tid == 0 && repeat > 1)
    printf("  %d MMAs (M=128 N=%d K=16 f16) in %lld cycles = %.1f cycles/MMA\n", repeat * 6, N, clock64() - t_start,
           double(clock64() - t_start) / (repeat * 6));
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(t0 + ((uint32_t)(warp * 32) << 16) + 64 + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<1>(t0, 512);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 128;
  std::vector<float> X(128 * K), W(N * K);
  srand(1);
  for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.3f) * 37.f;
  for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) * 0.21f;
  float ax = 0, aw = 0;
  for (auto v : X) ax = fmaxf(ax, fabsf(v));
  for (auto v : W) aw = fmaxf(aw, fabsf(v));
  const float xs = exp2f(13 - ceilf(log2f(ax))), ws = exp2f(13 - ceilf(log2f(aw)));
  std::vector<__half> whi(N * K), wlo(N * K);
  for (int i = 0; i < N * K; ++i) {
    float s = W[i] * ws;
    whi[i] = __float2half_rn(s);
    wlo[i] = __float2half_rn(s - __half2float(whi[i]));
  }
  float *dX, *dO; __half *dhi, *dlo;
  CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dO, 128 * N * 4)); CK(cudaMalloc(&dhi, N * K * 2)); CK(cudaMalloc(&dlo, N * K * 2));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dhi, whi.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dlo, wlo.data(), N * K * 2, cudaMemcpyHostToDevice));
  P p;
  PFN_encodeTiled enc = get_encode_tiled();
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {32, (cuuint32_t)N};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(i ? &p.blo : &p.bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, i ? (void*)dlo : (void*)dhi, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * N * 64 + 2048));
  for (int repeat : {1, 2000}) {
    probe<<<1, 128, 2 * N * 64 + 2048>>>(p, dX, xs, dO, N, repeat);
    CK(cudaDeviceSynchronize());
    if (repeat > 1) break;
    std::vector<float> O(128 * N);
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)X[m * K + k] * W[n * K + k];
        double got = (double)O[m * N + n] / ((double)xs * ws);
        maxerr = fmax(maxerr, fabs(got - ref));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("N=%d scales 2^%g 2^%g: max abs err %.3e  max |ref| %.3e  rel %.3e -> %s\n", N, log2(xs), log2(ws), maxerr, maxref,
           maxerr / maxref, maxerr / maxref < 3e-6 ? "OK" : "FAIL");
  }
  return 0;
}
