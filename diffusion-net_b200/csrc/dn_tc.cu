// tcgen05 tensor-core engine (sm_100a): the dense contractions of the DiffusionNetBlock path.
//
//   tc_rows_chain       fused chain of affine layers over 128-vertex row tiles
//                       (from_basis [+ complex-linear P|Q], MiniMLP + skip)      layers.py:56-67,229-239
//   tc_to_basis_partial split-V  Phi^T (M x)  with the 128x128 accumulator in TMEM   geometry.py:572-583
//
// Arithmetic: kind::tf32 MMAs with fp32 accumulation in TMEM.  "3x" mode splits every operand
// x = hi + lo (both exactly TF32) and issues lo*hi + hi*lo + hi*hi, recovering fp32-grade
// products; "1x" mode issues hi*hi only.
//
// Data movement: weights are pre-split and pre-laid-out in the UMMA canonical (no-swizzle,
// K-major) layout by a small pack kernel and streamed per K-chunk with bulk TMA copies
// (cp.async.bulk + mbarrier complete_tx).  Activations come from HBM (coalesced float4 loads)
// or from the previous layer's TMEM accumulator, are split in registers and stored straight
// into the canonical layout; a producer/consumer mbarrier ring hands K-chunks to the single
// MMA-issuing thread.
#include "dn_internal.h"
#include "dn_tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

using namespace tc;

constexpr int KC = 16;                          // k-elements per pipeline chunk (2 MMA k-steps of 8)
constexpr int TILE_M = 128;                     // vertex rows per tile == UMMA M
constexpr int NSA = 4, NSB = 4;                 // ring depths
constexpr int A_IMG = TILE_M * KC * 4;          // 8 KiB: one hi (or lo) A chunk image
constexpr int A_STAGE = 2 * A_IMG;              // hi + lo
constexpr int B_STAGE = 2 * 256 * KC * 4;       // 32 KiB: hi + lo at N = 256
constexpr int A_LBO = (TILE_M / 8) * 128;       // 2048 B between k-groups (4 elements) of A
constexpr int CHAIN_THREADS = 320;              // warp0 TMA, warp1 MMA, warps 2..9 workers (2 warpgroups)
constexpr int CHAIN_SMEM = NSA * A_STAGE + NSB * B_STAGE + 1024;

struct TcLayer {
  const float* wpack;
  const float* bias;
  const float* residual;
  int64_t ld_res;
  const float* row_scale;
  float* out;
  int64_t ld_out;
  int K, N, relu;
};

struct TcChainParams {
  DnRowsSrc src;
  TcLayer layer[DN_MAX_LAYERS];
  int n_layers;
  int passes;
  int variant;
  int64_t V;
};

// ---------------------------------------------------------------------------------------------
// weight pack:  W -> [chunk][hi | lo][ (k/4)*N*16B + (n/8)*128B + (n%8)*16B + (k%4)*4B ]
// ---------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ W, int64_t ldw, int w_trans, int K, int N,
                                    float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int n = idx / K, k = idx % K;
  const float w = w_trans ? W[(int64_t)k * ldw + n] : W[(int64_t)n * ldw + k];
  float hi, lo;
  split_tf32(w, hi, lo);
  const int chunk = k / KC, kk = k % KC;
  const int64_t img = (int64_t)N * KC;   // floats per image
  const int64_t off = (int64_t)chunk * 2 * img + (int64_t)(kk >> 2) * (N * 4) + (n >> 3) * 32 + (n & 7) * 4 + (kk & 3);
  dst[off] = hi;
  dst[off + img] = lo;
}

// ---------------------------------------------------------------------------------------------
// helpers shared by the worker warps
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split4(uint8_t* a_hi, uint8_t* a_lo, uint32_t byte_off, float4 v, int passes) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x);
  split_tf32(v.y, h.y, l.y);
  split_tf32(v.z, h.z, l.z);
  split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(a_hi + byte_off) = h;
  if (passes == 3) *reinterpret_cast<float4*>(a_lo + byte_off) = l;
}

// ---------------------------------------------------------------------------------------------
// fused affine chain over 128-row tiles
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CHAIN_THREADS, 1) rows_chain_kernel(const __grid_constant__ TcChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smA = smem;
  uint8_t* smB = smem + NSA * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSA * A_STAGE + NSB * B_STAGE);
  // bars: [0,NSA) a_full | [NSA,2NSA) a_empty | [2NSA, 2NSA+NSB) b_full | [.., +NSB) b_empty | d_full
  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + NSA);
  const uint32_t b_full = smem_u32(bars + 2 * NSA), b_empty = smem_u32(bars + 2 * NSA + NSB);
  const uint32_t d_full = smem_u32(bars + 2 * NSA + 2 * NSB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSA + 2 * NSB + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSA; ++i) { mbar_init(a_full + 8 * i, 4); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < NSB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int L = p.n_layers;
  const int64_t ntiles = (p.V + TILE_M - 1) / TILE_M;

  if (warp == 0) {
    // ===================== weight producer (bulk TMA) =====================
    if (lane == 0) {
      uint32_t ci = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int l = 0; l < L; ++l) {
          const int N = p.layer[l].N, nch = p.layer[l].K / KC;
          const uint32_t img_bytes = (uint32_t)N * KC * 4;
          const uint32_t bytes = p.passes == 3 ? 2 * img_bytes : img_bytes;
          for (int c = 0; c < nch; ++c, ++ci) {
            const uint32_t s = ci % NSB, ph = (ci / NSB) & 1;
            mbar_wait(b_empty + 8 * s, ph ^ 1);
            mbar_arrive_expect_tx(b_full + 8 * s, bytes);
            tma_bulk_g2s(smem_u32(smB + s * B_STAGE), p.layer[l].wpack + (int64_t)c * 2 * N * KC, bytes,
                         b_full + 8 * s);
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      uint32_t ci = 0, g = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int l = 0; l < L; ++l, ++g) {
          const int N = p.layer[l].N, nch = p.layer[l].K / KC;
          const uint32_t idesc = make_idesc_tf32(TILE_M, N);
          const uint32_t d_tmem = tmem_base + (g & 1) * 256;
          const uint32_t b_lbo = (uint32_t)N * 16, b_img = (uint32_t)N * KC * 4;
          uint32_t lboA = A_LBO, sboA = 128, lboB = b_lbo, sboB = 128;
          if (p.variant & 1) { lboA = 128; sboA = A_LBO; lboB = 128; sboB = b_lbo; }
          for (int c = 0; c < nch; ++c, ++ci) {
            const uint32_t sa = ci % NSA, pa = (ci / NSA) & 1, sb = ci % NSB, pb = (ci / NSB) & 1;
            mbar_wait(a_full + 8 * sa, pa);
            mbar_wait(b_full + 8 * sb, pb);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smA + sa * A_STAGE), a_lo = a_hi + A_IMG;
            const uint32_t b_hi = smem_u32(smB + sb * B_STAGE), b_lo = b_hi + b_img;
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
              const uint32_t ao = ks * 2 * A_LBO, bo = ks * 2 * b_lbo;
              const uint64_t dah = make_desc(a_hi + ao, lboA, sboA), dbh = make_desc(b_hi + bo, lboB, sboB);
              const uint32_t acc = (c | ks) ? 1u : 0u;
              if (p.passes == 3) {
                const uint64_t dal = make_desc(a_lo + ao, lboA, sboA), dbl = make_desc(b_lo + bo, lboB, sboB);
                mma_tf32_ss(d_tmem, dal, dbh, idesc, acc);
                mma_tf32_ss(d_tmem, dah, dbl, idesc, 1u);
                mma_tf32_ss(d_tmem, dah, dbh, idesc, 1u);
              } else {
                mma_tf32_ss(d_tmem, dah, dbh, idesc, acc);
              }
            }
            mma_commit(a_empty + 8 * sa);
            mma_commit(b_empty + 8 * sb);
          }
          mma_commit(d_full);
        }
    }
  } else {
    // ===================== workers: A-chunk producers + epilogue =====================
    const int wg = (warp - 2) >> 2;           // chunk parity this warpgroup owns
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int rl = lane & 7, kg = lane >> 3;  // conversion mapping: 8 rows x 4 k-groups per warp step
    uint32_t ci = 0, g = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t row0 = tile * TILE_M;
      for (int l = 0; l < L; ++l, ++g) {
        const TcLayer& Lr = p.layer[l];
        const int nch = Lr.K / KC;
        if (l == 0) {
          // ---- layer-0 operand from HBM: float4 loads -> hi/lo split -> canonical smem
          auto load_chunk = [&](int c, float4* r) {
            int k0 = c * KC, s = 0;
            while (s + 1 < p.src.nsrc && k0 >= p.src.width[s]) { k0 -= p.src.width[s]; ++s; }
            const float* base = p.src.ptr[s] + k0 + 4 * kg;
            const int64_t ld = p.src.ld[s];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int64_t row = row0 + 32 * quarter + 8 * it + rl;
              r[it] = (row < p.V) ? __ldg(reinterpret_cast<const float4*>(base + row * ld))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          };
          float4 cur[4], nxt[4];
          if (wg < nch) load_chunk(wg, cur);
          for (int c = wg; c < nch; c += 2) {
            if (c + 2 < nch) load_chunk(c + 2, nxt);
            const uint32_t cidx = ci + c, s = cidx % NSA, ph = (cidx / NSA) & 1;
            mbar_wait(a_empty + 8 * s, ph ^ 1);
            uint8_t* a_hi = smA + s * A_STAGE;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = 32 * quarter + 8 * it + rl;
              store_split4(a_hi, a_hi + A_IMG, kg * A_LBO + (r >> 3) * 128 + (r & 7) * 16, cur[it], p.passes);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * s);
#pragma unroll
            for (int it = 0; it < 4; ++it) cur[it] = nxt[it];
          }
        }
        const uint32_t ci_next = ci + nch;
        // ---- epilogue of layer l (and operand production for layer l+1)
        mbar_wait(d_full, g & 1);
        tc_fence_after();
        const bool has_next = (l + 1 < L);
        const int64_t row = row0 + 32 * quarter + lane;
        const int r = 32 * quarter + lane;
        const float rs = (Lr.row_scale && row < p.V) ? __ldg(Lr.row_scale + row) : 1.f;
        for (int c = wg; c < Lr.N / KC; c += 2) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(32 * quarter) << 16) + (g & 1) * 256 + c * KC, v);
          const int n0 = c * KC;
          if (Lr.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __ldg(Lr.bias + n0 + j);
          }
          if (Lr.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (Lr.row_scale) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= rs;
          }
          if (Lr.residual && row < p.V) {
            const float4* rp = reinterpret_cast<const float4*>(Lr.residual + row * Lr.ld_res + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 q = __ldg(rp + j);
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          }
          if (Lr.out && row < p.V) {
            float4* op = reinterpret_cast<float4*>(Lr.out + row * Lr.ld_out + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (has_next) {
            const uint32_t cidx = ci_next + c, s = cidx % NSA, ph = (cidx / NSA) & 1;
            mbar_wait(a_empty + 8 * s, ph ^ 1);
            uint8_t* a_hi = smA + s * A_STAGE;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              store_split4(a_hi, a_hi + A_IMG, j * A_LBO + (r >> 3) * 128 + (r & 7) * 16,
                           make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]), p.passes);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * s);
          }
        }
        tc_fence_before();
        ci = ci_next;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// to_basis, split over V:  partial[cta][k][c] = sum_{v in cta's range} Phi[v][k] * m[v] * x[v][c]
//   A = Phi^T (M = K_eig rows, padded to 128), B = (m x)^T (N = C rows); reduction dim = v.
//   Both operands are transposed on the fly: each lane loads a 4(v) x 4(k) block with float4
//   row loads and writes four 16-byte k-major vectors.  SBO is padded to 144 B so those
//   stores are bank-conflict free.
// ---------------------------------------------------------------------------------------------
constexpr int TB_SBO = 144;
constexpr int TB_LBO = 16 * TB_SBO;          // 16 eight-row groups (128 rows) per k-group
constexpr int TB_IMG = 4 * TB_LBO;           // 4 k-groups (16 v) : 9216 B
constexpr int TB_STAGE = 4 * TB_IMG;         // A hi, A lo, B hi, B lo
constexpr int TB_NS = 4;
constexpr int TB_THREADS = 320;              // warp0 idle/alloc, warp1 MMA, warps 2..9 workers
constexpr int TB_SMEM = TB_NS * TB_STAGE + 1024;

struct TcToBasisParams {
  const float* values;   // (V, C)
  const float* basis;    // (V, K)
  const float* mass;     // (V) or null
  float* partial;        // (grid, K, C)
  int64_t V;
  int K, C, passes, variant;
  int64_t chunks_per_cta;
};

__global__ void __launch_bounds__(TB_THREADS, 1) to_basis_kernel(const __grid_constant__ TcToBasisParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TB_NS * TB_STAGE);
  const uint32_t full = smem_u32(bars), empty = smem_u32(bars + TB_NS), d_full = smem_u32(bars + 2 * TB_NS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TB_NS + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < TB_NS; ++i) { mbar_init(full + 8 * i, 8); mbar_init(empty + 8 * i, 1); }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  // operand rows that no lane writes (k >= K or c >= C inside the 128-row images) must be zero
  for (int i = threadIdx.x; i < TB_NS * TB_STAGE / 16; i += blockDim.x)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  if (warp == 0) tmem_alloc<128>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t total_chunks = (p.V + KC - 1) / KC;
  const int64_t c_beg = (int64_t)blockIdx.x * p.chunks_per_cta;
  int64_t c_end = c_beg + p.chunks_per_cta;
  if (c_end > total_chunks) c_end = total_chunks;
  const int64_t nch = c_end > c_beg ? c_end - c_beg : 0;

  if (warp == 1) {
    if (lane == 0 && nch > 0) {
      const uint32_t idesc = make_idesc_tf32(128, p.C);
      uint32_t lbo = TB_LBO, sbo = TB_SBO;
      if (p.variant & 1) { lbo = TB_SBO; sbo = TB_LBO; }
      for (int64_t c = 0; c < nch; ++c) {
        const uint32_t s = c % TB_NS, ph = (c / TB_NS) & 1;
        mbar_wait(full + 8 * s, ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * TB_STAGE), a_lo = a_hi + TB_IMG, b_hi = a_hi + 2 * TB_IMG,
                       b_lo = a_hi + 3 * TB_IMG;
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {
          const uint32_t o = ks * 2 * TB_LBO;
          const uint64_t dah = make_desc(a_hi + o, lbo, sbo), dbh = make_desc(b_hi + o, lbo, sbo);
          const uint32_t acc = (c | ks) ? 1u : 0u;
          if (p.passes == 3) {
            const uint64_t dal = make_desc(a_lo + o, lbo, sbo), dbl = make_desc(b_lo + o, lbo, sbo);
            mma_tf32_ss(tmem_base, dal, dbh, idesc, acc);
            mma_tf32_ss(tmem_base, dah, dbl, idesc, 1u);
            mma_tf32_ss(tmem_base, dah, dbh, idesc, 1u);
          } else {
            mma_tf32_ss(tmem_base, dah, dbh, idesc, acc);
          }
        }
        mma_commit(empty + 8 * s);
      }
      mma_commit(d_full);
    }
  } else if (warp >= 2) {
    // warps 2..5: A = Phi^T, v-group (warp-2); warps 6..9: B = (m x)^T, v-group (warp-6)
    const int w = warp - 2;
    const bool isB = w >= 4;
    const int vg = w & 3;                           // which 4 of the chunk's 16 vertices
    const float* src = isB ? p.values : p.basis;
    const int width = isB ? p.C : p.K;
    const bool active = 4 * lane < width;           // this lane's 4 columns exist
    auto load4 = [&](int64_t chunk, float4* r) {
      const int64_t v0 = (c_beg + chunk) * KC + 4 * vg;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t v = v0 + j;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && v < p.V) {
          q = __ldg(reinterpret_cast<const float4*>(src + v * width + 4 * lane));
          if (isB && p.mass) {
            const float m = __ldg(p.mass + v);
            q.x *= m; q.y *= m; q.z *= m; q.w *= m;     // (values * massvec), geometry.py:583
          }
        }
        r[j] = q;
      }
    };
    float4 cur[4], nxt[4];
    if (nch > 0) load4(0, cur);
    for (int64_t c = 0; c < nch; ++c) {
      if (c + 1 < nch) load4(c + 1, nxt);
      const uint32_t s = c % TB_NS, ph = (c / TB_NS) & 1;
      mbar_wait(empty + 8 * s, ph ^ 1);
      if (active) {
        uint8_t* hi = smem + s * TB_STAGE + (isB ? 2 * TB_IMG : 0);
        uint8_t* lo = hi + TB_IMG;
        // transpose the 4(v) x 4(col) block: one 16-byte k-major vector per column
        const float col[4][4] = {{cur[0].x, cur[1].x, cur[2].x, cur[3].x},
                                 {cur[0].y, cur[1].y, cur[2].y, cur[3].y},
                                 {cur[0].z, cur[1].z, cur[2].z, cur[3].z},
                                 {cur[0].w, cur[1].w, cur[2].w, cur[3].w}};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = 4 * lane + i;   // operand row (eigen-index k, or channel c)
          store_split4(hi, lo, vg * TB_LBO + (m >> 3) * TB_SBO + (m & 7) * 16,
                       make_float4(col[i][0], col[i][1], col[i][2], col[i][3]), p.passes);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + 8 * s);
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
    }
    // ---- epilogue: TMEM -> partial[cta][k][c]   (warps 2..5 cover the four lane quarters)
    if (w < 4) {
      float* out = p.partial + (int64_t)blockIdx.x * p.K * p.C;
      const int quarter = warp & 3;
      const int k = 32 * quarter + lane;
      if (nch > 0) {
        mbar_wait(d_full, 0);
        tc_fence_after();
      }
      for (int c0 = 0; c0 < p.C; c0 += 16) {
        float v[16];
        if (nch > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(32 * quarter) << 16) + c0, v);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (k < p.K) {
          float4* op = reinterpret_cast<float4*>(out + (int64_t)k * p.C + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

int g_tc_ok = -1;
int g_sm_count = 0;

int env_variant() {
  const char* e = getenv("DN_TC_VARIANT");
  return e ? atoi(e) : 0;
}

}  // namespace

bool tc_supported_device() {
  if (g_tc_ok < 0) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      g_tc_ok = 0;
    } else {
      g_tc_ok = (prop.major == 10) ? 1 : 0;
      g_sm_count = prop.multiProcessorCount;
      if (g_tc_ok) {
        if (cudaFuncSetAttribute(rows_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM) !=
                cudaSuccess ||
            cudaFuncSetAttribute(to_basis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM) !=
                cudaSuccess) {
          cudaGetLastError();
          g_tc_ok = 0;
        }
      }
    }
    const char* off = getenv("DN_TC_DISABLE");
    if (off && atoi(off)) g_tc_ok = 0;
  }
  return g_tc_ok == 1;
}

int tc_rows_chain_supported(const DnRowsSrc& src, const DnLayer* layers, int n_layers) {
  if (n_layers < 1 || n_layers > DN_MAX_LAYERS) return DN_ERR_UNSUPPORTED;
  int k0 = 0;
  for (int s = 0; s < src.nsrc; ++s) {
    if (src.width[s] % KC || src.ld[s] % 4 || (reinterpret_cast<uintptr_t>(src.ptr[s]) & 15)) return DN_ERR_UNSUPPORTED;
    k0 += src.width[s];
  }
  if (k0 != layers[0].K) return DN_ERR_UNSUPPORTED;
  for (int l = 0; l < n_layers; ++l) {
    const DnLayer& L = layers[l];
    if (L.K % KC || L.K < 2 * KC || L.N % 16 || L.N < 16 || L.N > 256) return DN_ERR_UNSUPPORTED;
    if (L.emul || L.relu_mask_src) return DN_ERR_UNSUPPORTED;
    if (L.residual && (L.res_scale != 1.f || L.ld_res % 4 || (reinterpret_cast<uintptr_t>(L.residual) & 15)))
      return DN_ERR_UNSUPPORTED;
    if (L.out && (L.ld_out % 4 || (reinterpret_cast<uintptr_t>(L.out) & 15))) return DN_ERR_UNSUPPORTED;
    if (l > 0 && L.K != layers[l - 1].N) return DN_ERR_UNSUPPORTED;
  }
  if (!layers[n_layers - 1].out) return DN_ERR_UNSUPPORTED;
  return DN_OK;
}

int64_t tc_chain_ws_bytes(const DnLayer* layers, int n_layers) {
  int64_t b = 0;
  for (int l = 0; l < n_layers; ++l) b += ((int64_t)layers[l].K * layers[l].N * 2 * 4 + 255) / 256 * 256;
  return b;
}

int tc_rows_chain(const DnRowsSrc& src, const DnLayer* layers, int n_layers, int64_t V, int passes, void* ws,
                  int64_t ws_bytes, cudaStream_t st) {
  if (V <= 0) return DN_OK;
  if (!tc_supported_device()) return DN_ERR_NOT_SM100;
  if (tc_chain_ws_bytes(layers, n_layers) > ws_bytes || !ws) return DN_ERR_WORKSPACE;
  TcChainParams p;
  memset(&p, 0, sizeof(p));
  p.src = src;
  p.n_layers = n_layers;
  p.passes = passes;
  p.variant = env_variant();
  p.V = V;
  char* wp = static_cast<char*>(ws);
  for (int l = 0; l < n_layers; ++l) {
    const DnLayer& L = layers[l];
    float* dst = reinterpret_cast<float*>(wp);
    pack_weights_kernel<<<(L.K * L.N + 255) / 256, 256, 0, st>>>(L.W, L.ldw, L.w_trans, L.K, L.N, dst);
    DN_LAUNCH_CHECK();
    TcLayer& T = p.layer[l];
    T.wpack = dst; T.bias = L.bias; T.residual = L.residual; T.ld_res = L.ld_res; T.row_scale = L.row_scale;
    T.out = L.out; T.ld_out = L.ld_out; T.K = L.K; T.N = L.N; T.relu = L.relu;
    wp += ((int64_t)L.K * L.N * 2 * 4 + 255) / 256 * 256;
  }
  const int64_t ntiles = (V + TILE_M - 1) / TILE_M;
  const int grid = (int)(ntiles < g_sm_count ? ntiles : g_sm_count);
  rows_chain_kernel<<<grid, CHAIN_THREADS, CHAIN_SMEM, st>>>(p);
  DN_LAUNCH_CHECK();
  return DN_OK;
}

int tc_to_basis_supported(int K, int C) {
  if (K % 4 || K < 4 || K > 128) return DN_ERR_UNSUPPORTED;
  if (C % 16 || C < 16 || C > 128) return DN_ERR_UNSUPPORTED;
  return DN_OK;
}

int tc_to_basis_partial(const float* values, const float* basis, const float* massvec, int64_t V, int K, int C,
                        float* partial, int* P_out, int passes, cudaStream_t st) {
  if (!tc_supported_device()) return DN_ERR_NOT_SM100;
  if ((reinterpret_cast<uintptr_t>(values) & 15) || (reinterpret_cast<uintptr_t>(basis) & 15))
    return DN_ERR_UNSUPPORTED;
  TcToBasisParams p;
  p.values = values; p.basis = basis; p.mass = massvec; p.partial = partial;
  p.V = V; p.K = K; p.C = C; p.passes = passes; p.variant = env_variant();
  const int64_t total_chunks = (V + KC - 1) / KC;
  int grid = g_sm_count;
  if (total_chunks < grid) grid = (int)(total_chunks > 0 ? total_chunks : 1);
  p.chunks_per_cta = (total_chunks + grid - 1) / grid;
  if (p.chunks_per_cta < 1) p.chunks_per_cta = 1;
  grid = (int)((total_chunks + p.chunks_per_cta - 1) / p.chunks_per_cta);
  if (grid < 1) grid = 1;
  to_basis_kernel<<<grid, TB_THREADS, TB_SMEM, st>>>(p);
  DN_LAUNCH_CHECK();
  *P_out = grid;
  return DN_OK;
}
