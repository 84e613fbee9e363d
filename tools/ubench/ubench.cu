// Round-2 microbenchmarks that size the chain-kernel redesign (run on the GPU box, results in gpurun_out/ubench.log):
//   mma   : issue/execute rate of tcgen05.mma.kind::tf32 (M=128, N=128|256, K=8), A from TMEM or smem,
//           back to back, or through a minimal full/empty mbarrier ring with G MMAs per hand-off, 1 or 2 issuing warps
//   bulk  : L2 -> shared-memory bulk-copy (cp.async.bulk) bandwidth with every SM streaming the same / distinct buffers
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench.bin tools/ubench/ubench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include <algorithm>
#include "../../diffusion-net_b200/csrc/dn_tc_ptx.cuh"
using namespace tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct MmaP { int n, ts, G, nw, iters, ns; long long* cyc; };

// warps [0,nw): MMA issuers (own accumulator, own ring); warps [nw, 2nw): their zero-work producers
__global__ void __launch_bounds__(128, 1) mma_kernel(const MmaP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 64);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 48; ++i) mbar_init(smem_u32(bars + i), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(slot));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  const int NS = p.ns;
  if (warp < p.nw) {
    const uint32_t full = smem_u32(bars + warp * 16), empty = smem_u32(bars + warp * 16 + 8), done = smem_u32(bars + 40 + warp);
    const uint32_t idesc = make_idesc_tf32(128, p.n);
    const uint32_t d_tmem = tb + (p.nw == 2 ? warp * 128 : 0);          // 2 warps: N must be 128
    const uint32_t a_tmem = tb + 256 + warp * 64;
    const uint64_t a_desc = make_desc(smem_u32(smem) + warp * 8192, 2048, 128);
    const uint64_t b_desc = make_desc(smem_u32(smem) + 32768, (uint32_t)p.n * 16, 128);
    long long t0 = clock64();
    if (p.G == 0) {
      for (int i = 0; i < p.iters; ++i) {
        if (elect_one()) {
          if (p.ts) mma_tf32_ts(d_tmem, a_tmem, b_desc, idesc, 1u);
          else mma_tf32_ss(d_tmem, a_desc, b_desc, idesc, 1u);
        }
        __syncwarp();
      }
    } else {
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < p.iters; i += p.G) {
        mbar_wait(full + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          for (int g = 0; g < p.G; ++g) {
            if (p.ts) mma_tf32_ts(d_tmem, a_tmem, b_desc, idesc, 1u);
            else mma_tf32_ss(d_tmem, a_desc, b_desc, idesc, 1u);
          }
          mma_commit(empty + 8 * s);
        }
        __syncwarp();
        if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
      }
    }
    if (elect_one()) mma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) p.cyc[blockIdx.x * 2 + warp] = t1 - t0;
  } else if (warp < 2 * p.nw && p.G > 0) {
    const int w = warp - p.nw;
    const uint32_t full = smem_u32(bars + w * 16), empty = smem_u32(bars + w * 16 + 8);
    uint32_t s = 0, ph = 0;
    for (int i = 0; i < p.iters; i += p.G) {
      mbar_wait(empty + 8 * s, ph ^ 1);
      if ((threadIdx.x & 31) == 0) mbar_arrive(full + 8 * s);
      __syncwarp();
      if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<512>(tb);
}

struct BulkP { const uint8_t* src; long long per_cta_stride; int chunk, nchunks, iters; };
__global__ void __launch_bounds__(32, 1) bulk_kernel(const BulkP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * 32768);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(bars + i), 1); fence_barrier_init(); }
  __syncwarp();
  const uint8_t* src = p.src + (long long)blockIdx.x * p.per_cta_stride;
  if (threadIdx.x == 0) {
    uint32_t ph = 0;
    for (int i = 0; i < 4; ++i) {
      mbar_arrive_expect_tx(smem_u32(bars + i), p.chunk);
      tma_bulk_g2s(smem_u32(smem + i * 32768), src + (long long)(i % p.nchunks) * p.chunk, p.chunk, smem_u32(bars + i));
    }
    for (int i = 4; i < p.iters; ++i) {
      const int s = i & 3;
      if (s == 0 && i > 4) ph ^= 1;
      mbar_wait(smem_u32(bars + s), ((i - 4) >> 2) & 1);
      mbar_arrive_expect_tx(smem_u32(bars + s), p.chunk);
      tma_bulk_g2s(smem_u32(smem + s * 32768), src + (long long)(i % p.nchunks) * p.chunk, p.chunk, smem_u32(bars + s));
    }
    for (int i = p.iters; i < p.iters + 4; ++i) mbar_wait(smem_u32(bars + (i & 3)), ((i - 4) >> 2) & 1);
    (void)ph;
  }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  printf("SMs %d, max clock %d MHz\n", sms, khz / 1000);
  long long* cyc;
  CK(cudaMalloc(&cyc, sms * 2 * sizeof(long long)));
  const int SM = 96 * 1024 + 1024;
  CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct Cfg { int n, ts, G, nw, ns; };
  const Cfg cfgs[] = {{128, 1, 0, 1, 4}, {256, 1, 0, 1, 4}, {128, 0, 0, 1, 4}, {256, 0, 0, 1, 4},
                      {128, 1, 6, 1, 4}, {128, 1, 12, 1, 4}, {128, 1, 24, 1, 4}, {128, 1, 48, 1, 4},
                      {128, 1, 12, 1, 2}, {128, 1, 24, 1, 2}, {128, 1, 12, 1, 8},
                      {256, 1, 6, 1, 4}, {256, 1, 12, 1, 4}, {256, 1, 24, 1, 4},
                      {128, 1, 0, 2, 4}, {128, 1, 6, 2, 4}, {128, 1, 12, 2, 4}, {128, 1, 24, 2, 4}};
  for (const Cfg& c : cfgs) {
    MmaP p; p.n = c.n; p.ts = c.ts; p.G = c.G; p.nw = c.nw; p.ns = c.ns; p.iters = 4800; p.cyc = cyc;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(cyc, 0, sms * 2 * sizeof(long long)));
      CK(cudaEventRecord(e0));
      mma_kernel<<<sms, 128, SM>>>(p);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(sms * 2);
    CK(cudaMemcpy(h.data(), cyc, sms * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<long long> v;
    for (int i = 0; i < sms; ++i) for (int w = 0; w < c.nw; ++w) v.push_back(h[i * 2 + w]);
    std::sort(v.begin(), v.end());
    const double per = (double)v[v.size() / 2] / p.iters;
    const double flop = 2.0 * 128 * c.n * 8 * (double)p.iters * c.nw * sms;
    printf("mma N=%3d %s G=%2d stages=%d warps=%d : %.1f cyc/MMA per warp (min %.1f max %.1f)  kernel %.3f ms  => %.0f TFLOP/s tf32 chip\n",
           c.n, c.ts ? "TS" : "SS", c.G, c.ns, c.nw, per, (double)v.front() / p.iters, (double)v.back() / p.iters, ms,
           flop / (ms * 1e-3) / 1e12);
  }
  // ---- bulk copy bandwidth
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 64));
  const long long BUF = 640 * 1024;
  uint8_t* src;
  CK(cudaMalloc(&src, BUF * sms));
  CK(cudaMemset(src, 1, BUF * sms));
  for (int chunk : {32768, 16384}) for (int distinct = 0; distinct < 2; ++distinct) {
    BulkP p; p.src = src; p.per_cta_stride = distinct ? BUF : 0; p.chunk = chunk; p.nchunks = (int)(BUF / chunk); p.iters = 4000;
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      bulk_kernel<<<sms, 32, 4 * 32768 + 64>>>(p);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("bulk L2->smem chunk %5d %s: %.3f ms  => %.2f TB/s chip (%.1f B/clk/SM at max clock)\n", chunk,
           distinct ? "distinct 640KB per SM (93 MB, L2-resident)" : "one shared 640KB buffer", ms,
           (double)chunk * p.iters * sms / (ms * 1e-3) / 1e12, (double)chunk * p.iters / (ms * 1e-3) / (khz * 1e3));
  }
  return 0;
}
