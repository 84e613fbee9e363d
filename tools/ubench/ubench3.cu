// Round-2 microbenchmark: does the ~104-cycle cost of a 128x128x8 tcgen05.mma (one issuing warp, same accumulator every
// time) come from the accumulator dependency?  One warp issues back to back:
//   0: N=128, one accumulator            1: N=128, two accumulators alternating     2: N=64, halves [0,64) / [64,128) alternating
//   3: N=64, one accumulator             4: N=128, four accumulators round robin     5: N=128 tf32 / bf16(K=16) alternating, one acc
//   6: N=256, one accumulator            7: N=256, two accumulators alternating (A from shared memory)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench3.bin tools/ubench/ubench3.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include <algorithm>
#include "../../diffusion-net_b200/csrc/dn_tc_ptx.cuh"
using namespace tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct P { int mode, iters; long long* cyc; };

template <int MODE>
__global__ void __launch_bounds__(64, 1) k(const P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bars), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(slot));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (warp == 0) {
    const uint32_t done = smem_u32(bars);
    const uint32_t a_tmem = tb + 448;                     // operand columns [448, 512)
    const uint32_t i128 = make_idesc_tf32(128, 128), i64 = make_idesc_tf32(128, 64), i256 = make_idesc_tf32(128, 256);
    const uint32_t h128 = make_idesc_bf16(128, 128);
    const uint64_t a_desc = make_desc(smem_u32(smem), 2048, 128);
    const uint64_t b128 = make_desc(smem_u32(smem) + 32768, 128 * 16, 128), b64 = make_desc(smem_u32(smem) + 32768, 64 * 16, 128);
    const uint64_t b256 = make_desc(smem_u32(smem) + 32768, 256 * 16, 128);
    long long t0 = clock64();
    // (compile-time MODE and a 2x-unrolled warp-uniform loop: descriptors stay in uniform registers)
    for (int i = 0; i < p.iters; i += 2) {
      if (elect_one()) {
        if (MODE == 0) { mma_tf32_ts(tb, a_tmem, b128, i128, 1u); mma_tf32_ts(tb, a_tmem, b128, i128, 1u); }
        if (MODE == 1) { mma_tf32_ts(tb, a_tmem, b128, i128, 1u); mma_tf32_ts(tb + 128u, a_tmem, b128, i128, 1u); }
        if (MODE == 2) { mma_tf32_ts(tb, a_tmem, b64, i64, 1u); mma_tf32_ts(tb + 64u, a_tmem, b64, i64, 1u); }
        if (MODE == 3) { mma_tf32_ts(tb, a_tmem, b64, i64, 1u); mma_tf32_ts(tb, a_tmem, b64, i64, 1u); }
        if (MODE == 4) { mma_tf32_ts(tb, a_tmem, b128, i128, 1u); mma_tf32_ts(tb + 128u, a_tmem, b128, i128, 1u);
                         mma_tf32_ts(tb + 256u, a_tmem, b128, i128, 1u); mma_tf32_ts(tb, a_tmem, b128, i128, 1u); }   // 3 accumulators (4 MMAs)
        if (MODE == 5) { mma_tf32_ts(tb, a_tmem, b128, i128, 1u); mma_f16_ts(tb, a_tmem, b128, h128, 1u); }
        if (MODE == 6) { mma_tf32_ss(tb, a_desc, b256, i256, 1u); mma_tf32_ss(tb, a_desc, b256, i256, 1u); }
        if (MODE == 7) { mma_tf32_ss(tb, a_desc, b128, i128, 1u); mma_tf32_ss(tb + 256u, a_desc, b128, i128, 1u); }
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) p.cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<512>(tb);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* cyc;
  CK(cudaMalloc(&cyc, sizeof(long long) * sms));
  const char* names[] = {"N=128 one accumulator", "N=128 two accumulators alternating", "N=64 halves alternating", "N=64 one accumulator",
                         "N=128 three accumulators (4 MMAs per iteration)", "N=128 tf32 / bf16(K=16) alternating, one accumulator",
                         "N=256 SS one accumulator", "N=128 SS two accumulators alternating"};
  void (*kern[8])(const P) = {k<0>, k<1>, k<2>, k<3>, k<4>, k<5>, k<6>, k<7>};
  for (int mode = 0; mode < 8; ++mode) {
    CK(cudaFuncSetAttribute(kern[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    P p{mode, 4096, cyc};
    for (int rep = 0; rep < 2; ++rep) { kern[mode]<<<sms, 64, 100 * 1024>>>(p); CK(cudaDeviceSynchronize()); }
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    const int per_iter = mode == 4 ? 4 : 2;
    printf("mode %d %-52s: %.1f cycles per MMA (median over SMs)\n", mode, names[mode], (double)h[sms / 2] / (p.iters / 2) / per_iter);
  }
  return 0;
}
