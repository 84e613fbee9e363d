// Round-2 microbenchmark 2: how fast can one SM pull scattered rows (the sparse-gradient gather) through the TMA?
//   bulk  : cp.async.bulk of ROWB bytes per random row, issued by 1 or 32 lanes, NS slots of R rows in flight
//   g4    : cp.async.bulk.tensor.2d tile::gather4 (4 random rows x W floats per instruction), correctness + rate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/ubench2.bin tools/ubench/ubench2.cu
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include <string.h>
#include "../../diffusion-net_b200/csrc/dn_tc_ptx.cuh"
using namespace tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct P { const float* src; const int* rows; int nrows_per_cta; int rowb; int R; int lanes; float* sink; };

// slots: 2 x (R rows x rowb bytes); producer warp issues copies; consumer warp waits and releases immediately
__global__ void __launch_bounds__(64, 1) bulk_rows_kernel(const P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 98304);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(bars + i), 1); mbar_init(smem_u32(bars + 2 + i), 1); } fence_barrier_init(); }
  __syncthreads();
  const int* rows = p.rows + (long long)blockIdx.x * p.nrows_per_cta;
  const int nst = p.nrows_per_cta / p.R;
  if (warp == 0) {
    for (int s = 0; s < nst; ++s) {
      const int b = s & 1;
      mbar_wait(smem_u32(bars + 2 + b), ((s >> 1) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(smem_u32(bars + b), (uint32_t)(p.R * p.rowb));
      __syncwarp();
      if (p.lanes == 1) {
        if (lane == 0)
          for (int r = 0; r < p.R; ++r)
            tma_bulk_g2s(smem_u32(smem + b * 98304 + r * p.rowb), p.src + (long long)rows[s * p.R + r] * (p.rowb / 4), p.rowb, smem_u32(bars + b));
      } else {
        for (int r = lane; r < p.R; r += 32)
          tma_bulk_g2s(smem_u32(smem + b * 98304 + r * p.rowb), p.src + (long long)rows[s * p.R + r] * (p.rowb / 4), p.rowb, smem_u32(bars + b));
      }
      __syncwarp();
    }
  } else {
    float acc = 0.f;
    for (int s = 0; s < nst; ++s) {
      const int b = s & 1;
      mbar_wait(smem_u32(bars + b), (s >> 1) & 1);
      acc += reinterpret_cast<float*>(smem + b * 98304)[lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bars + 2 + b));
    }
    if (acc == 123.456f) p.sink[0] = acc;
  }
}

struct G { CUtensorMap map; const int* rows; int nrows_per_cta; int W; int R; float* out; int check; };
__global__ void __launch_bounds__(64, 1) g4_kernel(const __grid_constant__ G p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 98304);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(bars + i), 1); mbar_init(smem_u32(bars + 2 + i), 1); } fence_barrier_init(); }
  __syncthreads();
  const int* rows = p.rows + (long long)blockIdx.x * p.nrows_per_cta;
  const int nst = p.nrows_per_cta / p.R;
  const int rowb = p.W * 4;
  if (warp == 0) {
    for (int s = 0; s < nst; ++s) {
      const int b = s & 1;
      mbar_wait(smem_u32(bars + 2 + b), ((s >> 1) & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(smem_u32(bars + b), (uint32_t)(p.R * rowb));
        for (int r = 0; r < p.R; r += 4) {
          const int* q = rows + s * p.R + r;
          asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                       :: "r"(smem_u32(smem + b * 98304 + r * rowb)), "l"(reinterpret_cast<uint64_t>(&p.map)), "r"(0), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]),
                          "r"(smem_u32(bars + b)) : "memory");
        }
      }
      __syncwarp();
    }
  } else {
    float acc = 0.f;
    for (int s = 0; s < nst; ++s) {
      const int b = s & 1;
      mbar_wait(smem_u32(bars + b), (s >> 1) & 1);
      if (p.check && blockIdx.x == 0 && s == 0) {
        // row r of the stage should hold the source row rows[r]: element j = rows[r] * 1000 + j (see main)
        for (int r = 0; r < 8; ++r) p.out[r * 32 + lane] = reinterpret_cast<float*>(smem + r * rowb)[lane * (p.W / 32)];
      }
      acc += reinterpret_cast<float*>(smem + b * 98304)[lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bars + 2 + b));
    }
    if (acc == 123.456f) p.out[0] = acc;
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const bool do_bulk = argc < 2 || !strcmp(argv[1], "bulk");
  const bool do_g4 = argc >= 4 && !strcmp(argv[1], "g4");
  const int g4_w = do_g4 ? atoi(argv[2]) : 0, g4_rows = do_g4 ? atoi(argv[3]) : 0;
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const int V = 200000, W = 384;
  float* src; CK(cudaMalloc(&src, (size_t)V * W * 4));
  std::vector<float> h((size_t)V * W);
  for (int v = 0; v < V; ++v) for (int j = 0; j < W; ++j) h[(size_t)v * W + j] = (float)(v % 10000) * 1000.f + j;
  CK(cudaMemcpy(src, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  const int per_cta = 6144;                      // rows fetched by each CTA
  std::vector<int> hr((size_t)sms * per_cta);
  // mesh-like locality: each CTA walks a window of the vertex range, rows within +-600 of a moving centre
  srand(1);
  for (int c = 0; c < sms; ++c) for (int i = 0; i < per_cta; ++i) {
    int centre = (int)(((long long)c * per_cta + i) * (long long)V / ((long long)sms * per_cta));
    int r = centre + (rand() % 1201) - 600; if (r < 0) r = 0; if (r >= V) r = V - 1;
    hr[(size_t)c * per_cta + i] = r;
  }
  int* rows; CK(cudaMalloc(&rows, hr.size() * 4)); CK(cudaMemcpy(rows, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
  float* sink; CK(cudaMalloc(&sink, 4096));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int SM = 2 * 98304 + 64;
  CK(cudaFuncSetAttribute(bulk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  CK(cudaFuncSetAttribute(g4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  if (do_bulk) for (int rowb : {1536, 1024, 512}) for (int lanes : {1, 32}) for (int R : {32, 64}) {
    if (R * rowb > 98304) continue;
    P p; p.src = src; p.rows = rows; p.nrows_per_cta = per_cta; p.rowb = rowb; p.R = R; p.lanes = lanes; p.sink = sink;
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) { CK(cudaEventRecord(e0)); bulk_rows_kernel<<<sms, 64, SM>>>(p); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, e0, e1)); }
    // NOTE rows are W*4 = 1536 B apart; a rowb < 1536 copy reads the head of the row
    printf("bulk rows %4d B x %2d per stage, %2d issuing lane(s): %.3f ms => %.2f TB/s chip, %.1f B/clk/SM, %.0f cycles per copy\n", rowb, R, lanes, ms,
           (double)rowb * per_cta * sms / (ms * 1e-3) / 1e12, (double)rowb * per_cta / (ms * 1e-3) / (khz * 1e3), (ms * 1e-3) * (khz * 1e3) / per_cta);
  }
  // ---- gather4
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncFn enc = (EncFn)f;
  float* out; CK(cudaMalloc(&out, 4096)); 
  if (do_g4) for (int Wb : {g4_w}) for (int boxrows : {g4_rows}) {
    G g; memset(&g, 0, sizeof(g));
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)V};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {(cuuint32_t)Wb, (cuuint32_t)boxrows};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&g.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("g4 W=%d boxrows=%d: encode failed %d\n", Wb, boxrows, (int)r); continue; }
    g.rows = rows; g.nrows_per_cta = per_cta; g.W = Wb; g.R = 32; g.out = out; g.check = 1;
    CK(cudaMemset(out, 0, 4096));
    float ms = 0; cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 3 && err == cudaSuccess; ++rep) {
      CK(cudaEventRecord(e0)); g4_kernel<<<sms, 64, SM>>>(g); CK(cudaEventRecord(e1));
      err = cudaDeviceSynchronize();
      if (err == cudaSuccess) CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    if (err != cudaSuccess) { printf("g4 W=%d boxrows=%d: kernel failed: %s\n", Wb, boxrows, cudaGetErrorString(err)); return 0; }
    std::vector<float> ho(1024); CK(cudaMemcpy(ho.data(), out, 4096, cudaMemcpyDeviceToHost));
    int ok = 1;
    for (int r8 = 0; r8 < 8; ++r8) { float want = (float)(hr[r8] % 10000) * 1000.f + 0; if (ho[r8 * 32] != want) ok = 0; }
    printf("g4 W=%3d floats, tensor-map box rows %d: %s (row0 got %.0f want %.0f; row5 got %.0f want %.0f)  %.3f ms => %.2f TB/s chip, %.1f B/clk/SM, %.0f cycles per gather4\n",
           Wb, boxrows, ok ? "DATA OK" : "DATA MISMATCH", ho[0], (float)(hr[0] % 10000) * 1000.f, ho[5 * 32], (float)(hr[5] % 10000) * 1000.f, ms,
           (double)Wb * 4 * per_cta * sms / (ms * 1e-3) / 1e12, (double)Wb * 4 * per_cta / (ms * 1e-3) / (khz * 1e3), (ms * 1e-3) * (khz * 1e3) / (per_cta / 4));
  }
  return 0;
}
