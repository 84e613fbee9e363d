// bf16 fused affine chain over 128-vertex row tiles (sm_100a) -- the dense layers of the DiffusionNetBlock path in
// DN_ENGINE_BF16 (BASELINE config 3: C_width = 256, bf16):
//     from_basis, the commuted complex-linear P and Q                   reference layers.py:64-67, 117-126
//     cat(x_in, x_diffuse, features) -> MiniMLP -> + x_in              reference layers.py:133-164, 229-239
//
// Same role layout as rows_chain3_kernel (dn_chain.cu); what differs is the arithmetic and where the operands live:
//   * one tcgen05.mma.kind::f16 pass (bf16 x bf16 -> fp32), K = 16 per instruction, four MMAs per 64-wide K-stage;
//   * layers up to 256 wide CHAIN: the two 256-column accumulators ping-pong over all 512 TMEM columns, so the A operand
//     cannot live in tensor memory -- it is written as bf16 in the UMMA canonical K-major layout into a two-slot
//     shared-memory ring (16 KiB per stage) and the MMAs read both operands through shared-memory descriptors;
//   * fp32 in HBM on both sides (module API unchanged): layer-0 rows arrive as fp32 TMA boxes and are rounded to bf16
//     on the way into the operand ring; outputs are fp32.
//
// One persistent CTA per SM, 16 warps:
//   warp 0      weight producer: one bulk copy per 64-wide K-stage of pre-packed bf16 weights (3-slot ring, 32 KiB slots)
//   warp 1      row-box producer: layer-0 operands as 128-row x 32-column fp32 boxes (SWIZZLE_128B), two boxes per stage;
//               each operand warpgroup owns two of the four slots
//   warp 2      MMA issuer
//   warp 3      allocates / frees TMEM
//   warps 4-11  two operand warpgroups; stage i (counted over the whole launch) belongs to warpgroup i & 1 and to
//               operand slot i & 1, so every ring slot has one in-order producer and one in-order consumer
//   warps 12-15 output warpgroup (accumulator -> bias / ReLU / row scale / residual or ReLU mask -> staging -> TMA store)
//
// Shared memory: row boxes 4 x 16 KiB | operand ring 2 x 16 KiB | output staging 2 x 16 KiB | weight ring 3 x 32 KiB.
// Envelope (tc_chain16_supported): every K % 64 == 0, layer-0 sources % 64, every N % 32 == 0 (% 64 when it feeds
// another layer), N <= 256.
#include "dn_internal.h"
#include "dn_tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace {

using namespace tc;

constexpr int H_THREADS = 512;
constexpr int H_KS = 64;                        // k-columns per pipeline stage
constexpr int H_TILE = 128;                     // rows per tile == UMMA M
constexpr int H_BOX = H_TILE * 32 * 4;          // 16 KiB fp32 row box (32 columns)
constexpr int H_ASTAGE = H_TILE * H_KS * 2;     // 16 KiB bf16 operand stage
constexpr int H_WSLOT = 256 * H_KS * 2;         // 32 KiB weight slot (N <= 256)
constexpr int H_NW = 3;
constexpr int H_OFF_A = 4 * H_BOX;
constexpr int H_OFF_OUT = H_OFF_A + 2 * H_ASTAGE;
constexpr int H_OFF_W = H_OFF_OUT + 2 * H_BOX;
constexpr int H_OFF_BAR = H_OFF_W + H_NW * H_WSLOT;
constexpr int H_SMEM = H_OFF_BAR + 512;
constexpr int H_A_LBO = (H_TILE / 8) * 128;     // 2048 B between k-groups (8 bf16) of the A operand

struct HLayer {
  const void* wpack;       // pack_fmt 2: 64-wide stages of bf16, K-major canonical
  const float* bias;
  const float* row_scale;
  int K, N, relu;
  int has_out;
  int has_res;             // 0 | 1: + residual | 2: * (aux > 0)
};

struct HParams {
  HLayer layer[DN_MAX_LAYERS];
  int n_layers, nsrc;
  int src_width[DN_MAX_SRC];
  int64_t V;
  const int32_t* tile_group;   // optional (mesh batches): layer 0 of tile t streams packed matrix tile_group[t]
  int64_t group_stride;        //   (floats between the per-mesh matrices)
};

struct HMaps {
  CUtensorMap src[DN_MAX_SRC];
  CUtensorMap out[DN_MAX_LAYERS];
  CUtensorMap res;
};

__device__ __forceinline__ void h_box_load(uint32_t dst_smem, const CUtensorMap* tmap, int col, int row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col), "r"(row), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void h_box_store(const CUtensorMap* tmap, int col, int row, uint32_t src_smem) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(col), "r"(row), "r"(src_smem)
               : "memory");
}
__device__ __forceinline__ void h_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void h_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void h_prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void h_tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float4 h_lds128(uint32_t a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ void h_sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void h_sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate (K = 16)
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(H_THREADS, 1)
rows_chain16_kernel(const __grid_constant__ HParams p, const __grid_constant__ HMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem0 = smem_u32(smem);
  if (smem0 & 1023u) __trap();
  const uint32_t raw_u = smem0, a_u = smem0 + H_OFF_A, out_u = smem0 + H_OFF_OUT, w_u = smem0 + H_OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + H_OFF_BAR);
  // bars: raw_full[2] raw_empty[2] a_full[2] a_empty[2] w_full[3] w_empty[3] dm_full[2] do_full[2] dm_empty[2] do_empty[2]
  //       res[4 warps][2] done
  //   raw_full / raw_empty [wg]: the two row boxes of warpgroup wg's next layer-0 stage
  //   a_full / a_empty [s]: operand slot s (4 warp arrivals / one tcgen05.commit)
  //   w_full / w_empty [s]: weight slot s (expect_tx / one tcgen05.commit)
  const uint32_t raw_full = smem_u32(bars), raw_empty = smem_u32(bars + 2);
  const uint32_t a_full = smem_u32(bars + 4), a_empty = smem_u32(bars + 6);
  const uint32_t w_full = smem_u32(bars + 8), w_empty = smem_u32(bars + 11);
  const uint32_t dm_full = smem_u32(bars + 14), do_full = smem_u32(bars + 16);
  const uint32_t dm_empty = smem_u32(bars + 18), do_empty = smem_u32(bars + 20);
  const uint32_t res_bar = smem_u32(bars + 22), done_bar = smem_u32(bars + 30);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(raw_full + 8 * i, 1); mbar_init(raw_empty + 8 * i, 4);
      mbar_init(a_full + 8 * i, 4);   mbar_init(a_empty + 8 * i, 1);
      mbar_init(dm_full + 8 * i, 1);  mbar_init(do_full + 8 * i, 1);
      mbar_init(dm_empty + 8 * i, 8); mbar_init(do_empty + 8 * i, 4);
    }
    for (int i = 0; i < H_NW; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < 8; ++i) mbar_init(res_bar + 8 * i, 1);
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 3) tmem_alloc<512>(smem_u32(tmem_slot));
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) h_prefetch_tmap(&maps.src[s]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int L = p.n_layers;
  const int64_t ntiles = (p.V + H_TILE - 1) / H_TILE;
  const int nst0 = p.layer[0].K / H_KS;
  int S = 0;                                                     // stages per tile
  for (int l = 0; l < L; ++l) S += p.layer[l].K / H_KS;

  if (warp == 0) {
    // ===================== weight producer =====================
    uint32_t i = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int l = 0; l < L; ++l) {
        const int N = p.layer[l].N, nst = p.layer[l].K / H_KS;
        const uint32_t bytes = (uint32_t)N * (H_KS * 2);
        const char* wsrc = static_cast<const char*>(p.layer[l].wpack);
        if (l == 0 && p.tile_group) wsrc += (int64_t)__ldg(p.tile_group + tile) * p.group_stride * 4;
        for (int c = 0; c < nst; ++c, ++i) {
          const uint32_t s = i % H_NW;
          mbar_wait(w_empty + 8 * s, ((i / H_NW) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(w_full + 8 * s, bytes);
            tma_bulk_g2s(w_u + s * H_WSLOT, wsrc + (int64_t)c * bytes, bytes, w_full + 8 * s);
          }
          __syncwarp();
        }
      }
  } else if (warp == 1) {
    // ===================== row-box producer (layer-0 operands) =====================
    uint32_t cnt0 = 0, cnt1 = 0;                                 // layer-0 stages handed to warpgroup 0 / 1 so far
    uint32_t t_seq = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_seq) {
      int s = 0, k0 = 0;
      const uint32_t i_base = t_seq * (uint32_t)S;
      for (int c = 0; c < nst0; ++c) {
        const uint32_t wg = (i_base + (uint32_t)c) & 1u;
        const uint32_t n = wg ? cnt1 : cnt0;
        if (wg) ++cnt1; else ++cnt0;
        mbar_wait(raw_empty + 8 * wg, (n & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(raw_full + 8 * wg, 2 * H_BOX);
          h_box_load(raw_u + (2 * wg) * H_BOX, &maps.src[s], k0, (int)(tile * H_TILE), raw_full + 8 * wg);
          h_box_load(raw_u + (2 * wg + 1) * H_BOX, &maps.src[s], k0 + 32, (int)(tile * H_TILE), raw_full + 8 * wg);
        }
        __syncwarp();
        k0 += H_KS;
        if (k0 == p.src_width[s]) { k0 = 0; ++s; }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    uint32_t i = 0, g = 0;
    uint32_t cm0 = 0, cm1 = 0, co0 = 0, co1 = 0;
    uint32_t pend = 0;
    const uint64_t tmplA = make_desc(0, H_A_LBO, 128);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int l = 0; l < L; ++l, ++g) {
        const int N = p.layer[l].N, nst = p.layer[l].K / H_KS;
        const uint32_t idesc = make_idesc_bf16(H_TILE, N);
        const uint32_t buf = g & 1u;
        const uint32_t d_tmem = tmem_base + buf * 256u;
        const uint32_t b_lbo = (uint32_t)N * 16u;
        const uint64_t tmplB = make_desc(0, b_lbo, 128);
        if (pend & (1u << buf)) {
          mbar_wait(dm_empty + 8 * buf, ((buf ? cm1 : cm0) - 1u) & 1u);
          pend &= ~(1u << buf);
        }
        if (pend & (4u << buf)) {
          mbar_wait(do_empty + 8 * buf, ((buf ? co1 : co0) - 1u) & 1u);
          pend &= ~(4u << buf);
        }
        tc_fence_after();
        for (int c = 0; c < nst; ++c, ++i) {
          const uint32_t sa = i & 1u, sw = i % H_NW;
          mbar_wait(a_full + 8 * sa, (i >> 1) & 1u);
          mbar_wait(w_full + 8 * sw, (i / H_NW) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = tmplA + ((a_u + sa * H_ASTAGE) >> 4);
            const uint64_t db = tmplB + ((w_u + sw * H_WSLOT) >> 4);
#pragma unroll
            for (int j = 0; j < H_KS / 16; ++j)
              mma_f16_ss(d_tmem, da + (uint32_t)j * ((2u * H_A_LBO) >> 4), db + (uint32_t)j * ((2u * b_lbo) >> 4), idesc,
                         (c | j) ? 1u : 0u);
            mma_commit(a_empty + 8 * sa);
            mma_commit(w_empty + 8 * sw);
            if (c + 1 == nst) {
              if (l + 1 < L) mma_commit(dm_full + 8 * buf);
              if (p.layer[l].has_out) mma_commit(do_full + 8 * buf);
            }
          }
          __syncwarp();
        }
        if (l + 1 < L) { pend |= (1u << buf); if (buf) ++cm1; else ++cm0; }
        if (p.layer[l].has_out) { pend |= (4u << buf); if (buf) ++co1; else ++co0; }
      }
    if (elect_one()) mma_commit(done_bar);
    __syncwarp();
    mbar_wait(done_bar, 0);
  } else if (warp >= 4 && warp < 12) {
    // ===================== operand warpgroups: layer-0 conversion + chained epilogues =====================
    const uint32_t wg = (uint32_t)(warp - 4) >> 2;
    const int quarter = warp & 3;
    const int trow = 32 * quarter + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * quarter) << 16);
    const uint32_t swz = (uint32_t)(trow & 7);
    const uint32_t a_row = a_u + wg * H_ASTAGE + (uint32_t)(trow >> 3) * 128u + (uint32_t)(trow & 7) * 16u;   // slot == wg
    uint32_t nraw = 0;                                           // layer-0 stages this warpgroup has taken
    uint32_t nput = 0;                                           // operand stages this warpgroup has produced

    // 32 fp32 -> k-groups [kg0, kg0 + 4) of this warpgroup's operand slot
    auto put_half = [&](const float* x, int kg0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        h_sts128u(a_row + (uint32_t)(kg0 + q) * H_A_LBO, pack_bf16x2(x[8 * q], x[8 * q + 1]),
                  pack_bf16x2(x[8 * q + 2], x[8 * q + 3]), pack_bf16x2(x[8 * q + 4], x[8 * q + 5]),
                  pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
    };
    auto slot_acquire = [&]() {                                  // my slot's previous stage has been consumed by the MMAs
      mbar_wait(a_empty + 8 * wg, (nput & 1u) ^ 1u);
      tc_fence_after();
    };
    auto slot_publish = [&]() {
      fence_proxy_async();                                       // generic-proxy writes -> visible to the UMMA reads
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8 * wg);
      ++nput;
    };

    uint32_t t_seq = 0;
    uint32_t um0 = 0, um1 = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_seq) {
      const uint32_t i_base = t_seq * (uint32_t)S;
      // ---- layer 0: my stages of the row boxes -> bf16 operand slot
      for (int c = (int)((wg ^ i_base) & 1u); c < nst0; c += 2) {
        mbar_wait(raw_full + 8 * wg, nraw & 1u);
        ++nraw;
        float x[64];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t rowaddr = raw_u + (2 * wg + (uint32_t)h) * H_BOX + (uint32_t)trow * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q = h_lds128(rowaddr + ((((uint32_t)j) ^ swz) << 4));
            x[32 * h + 4 * j] = q.x; x[32 * h + 4 * j + 1] = q.y; x[32 * h + 4 * j + 2] = q.z; x[32 * h + 4 * j + 3] = q.w;
          }
        }
        // the slot release must not overtake the loads: really consume one register of each 16-byte load first
        uint32_t d = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) d ^= __float_as_uint(x[4 * j]);
        consume_loaded(d);
        __syncwarp();
        if (lane == 0) mbar_arrive(raw_empty + 8 * wg);
        slot_acquire();
        put_half(x, 0);
        put_half(x + 32, 4);
        slot_publish();
      }
      // ---- chained layers: accumulator l -> operand stages of layer l + 1
      uint32_t ib = i_base + (uint32_t)nst0;
      for (int l = 0; l + 1 < L; ++l) {
        const HLayer& Lr = p.layer[l];
        const uint32_t g = t_seq * (uint32_t)L + (uint32_t)l;
        const uint32_t buf = g & 1u;
        const int nco = Lr.N / H_KS;                             // <= 4
        const int c_first = (int)((wg ^ ib) & 1u);
        float bl[2][2] = {{0.f, 0.f}, {0.f, 0.f}};               // bias lanes of my (at most two) chunks
        if (Lr.bias) {
#pragma unroll
          for (int q = 0; q < 2; ++q)
            if (c_first + 2 * q < nco) {
              bl[q][0] = __ldg(Lr.bias + (c_first + 2 * q) * H_KS + lane);
              bl[q][1] = __ldg(Lr.bias + (c_first + 2 * q) * H_KS + 32 + lane);
            }
        }
        mbar_wait(dm_full + 8 * buf, (buf ? um1 : um0) & 1u);
        if (buf) ++um1; else ++um0;
        tc_fence_after();
        const uint32_t d_lane = lane_base + buf * 256u;
        bool released = false;
        for (int c = c_first; c < nco; c += 2) {
          float v[64];
          h_tmem_ld32(d_lane + (uint32_t)c * H_KS, v);
          h_tmem_ld32(d_lane + (uint32_t)c * H_KS + 32, v + 32);
          if (c + 2 >= nco) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dm_empty + 8 * buf);
            released = true;
          }
          if (Lr.bias) {
            const int qi = (c - c_first) >> 1;
            const float m0 = qi == 0 ? bl[0][0] : bl[1][0], m1 = qi == 0 ? bl[0][1] : bl[1][1];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] += __shfl_sync(0xffffffffu, m0, j);
              v[32 + j] += __shfl_sync(0xffffffffu, m1, j);
            }
          }
          if (Lr.relu) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          slot_acquire();
          put_half(v, 0);
          put_half(v + 32, 4);
          slot_publish();
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dm_empty + 8 * buf);
        }
        ib += (uint32_t)nco;
      }
    }
  } else if (warp >= 12) {
    // ===================== output warpgroup =====================
    const int quarter = warp & 3;
    const int trow = 32 * quarter + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * quarter) << 16);
    const uint32_t swz = (uint32_t)(lane & 7);
    const uint32_t my_res = res_bar + 8u * (uint32_t)(quarter * 2);
    uint32_t oc = 0;
    uint32_t rcbits = 0;
    uint32_t uo0 = 0, uo1 = 0;
    uint32_t t_seq = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_seq) {
      const int row0 = (int)(tile * H_TILE) + 32 * quarter;
      const int64_t row = tile * H_TILE + trow;
      for (int l = 0; l < L; ++l) {
        const HLayer& Lr = p.layer[l];
        if (!Lr.has_out) continue;
        const uint32_t g = t_seq * (uint32_t)L + (uint32_t)l;
        const uint32_t buf = g & 1u;
        const uint32_t use = buf ? uo1 : uo0;
        if (buf) ++uo1; else ++uo0;
        const int nco = Lr.N / 32;
        const bool res = Lr.has_res != 0;
        const float rs = (Lr.row_scale && row < p.V) ? __ldg(Lr.row_scale + row) : 1.f;
        const uint32_t d_lane = lane_base + buf * 256u;
        float bl[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (Lr.bias) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < nco) bl[q] = __ldg(Lr.bias + q * 32 + lane);
        }
        if (!res) {
          mbar_wait(do_full + 8 * buf, use & 1u);
          tc_fence_after();
        }
        for (int c0 = 0; c0 < nco; c0 += 2) {
          const int ng = (nco - c0) < 2 ? (nco - c0) : 2;
          if (res) {
            if (lane == 0) {
              h_bulk_wait_read<0>();
              for (int j = 0; j < ng; ++j) {
                const uint32_t b = (oc + (uint32_t)j) & 1u;
                mbar_arrive_expect_tx(my_res + 8 * b, 4096);
                h_box_load(out_u + b * H_BOX + (uint32_t)quarter * 4096u, &maps.res, (c0 + j) * 32, row0, my_res + 8 * b);
              }
            }
            __syncwarp();
            if (c0 == 0) {
              mbar_wait(do_full + 8 * buf, use & 1u);
              tc_fence_after();
            }
          }
          for (int j = 0; j < ng; ++j, ++oc) {
            const int c = c0 + j;
            const uint32_t b = oc & 1u;
            const uint32_t slice = out_u + b * H_BOX + (uint32_t)quarter * 4096u;
            if (!res) {
              if (lane == 0) h_bulk_wait_read<1>();
              __syncwarp();
            }
            float v[32];
            h_tmem_ld32(d_lane + (uint32_t)c * 32, v);
            if (c + 1 == nco) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(do_empty + 8 * buf);
            }
            if (Lr.bias) {
              float mine = bl[0];
#pragma unroll
              for (int q = 1; q < 8; ++q) mine = (c == q) ? bl[q] : mine;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] += __shfl_sync(0xffffffffu, mine, jj);
            }
            if (Lr.relu) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] = fmaxf(v[jj], 0.f);
            }
            if (Lr.row_scale) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] *= rs;
            }
            const uint32_t rowaddr = slice + (uint32_t)lane * 128u;
            if (res) {
              mbar_wait(my_res + 8 * b, (rcbits >> b) & 1u);
              rcbits ^= (1u << b);
              if (Lr.has_res == 1) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const float4 q = h_lds128(rowaddr + ((((uint32_t)jj) ^ swz) << 4));
                  v[4 * jj] += q.x; v[4 * jj + 1] += q.y; v[4 * jj + 2] += q.z; v[4 * jj + 3] += q.w;
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const float4 q = h_lds128(rowaddr + ((((uint32_t)jj) ^ swz) << 4));
                  v[4 * jj] = q.x > 0.f ? v[4 * jj] : 0.f;         v[4 * jj + 1] = q.y > 0.f ? v[4 * jj + 1] : 0.f;
                  v[4 * jj + 2] = q.z > 0.f ? v[4 * jj + 2] : 0.f; v[4 * jj + 3] = q.w > 0.f ? v[4 * jj + 3] : 0.f;
                }
              }
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              h_sts128(rowaddr + ((((uint32_t)jj) ^ swz) << 4), v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              h_box_store(&maps.out[l], c * 32, row0, slice);
              h_bulk_commit();
            }
            __syncwarp();
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 3) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmaEncodeFn h_encode_fn() {
  static TmaEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmaEncodeFn>(f);
    else
      cudaGetLastError();
  }
  return fn;
}

int h_box_map(CUtensorMap* m, const float* ptr, int width, int64_t ld, int64_t V, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)V};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = h_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

constexpr int kMaxDev = 64;
int g_attr_dev[kMaxDev];

}  // namespace

int tc_chain16_supported(const DnRowsSrc& src, const DnLayer* layers, int n_layers) {
  if (h_encode_fn() == nullptr) return DN_ERR_UNSUPPORTED;
  if (n_layers < 1 || n_layers > DN_MAX_LAYERS) return DN_ERR_UNSUPPORTED;
  int k0 = 0;
  for (int s = 0; s < src.nsrc; ++s) {
    if (src.width[s] % H_KS || src.ld[s] % 4 || (reinterpret_cast<uintptr_t>(src.ptr[s]) & 15)) return DN_ERR_UNSUPPORTED;
    k0 += src.width[s];
  }
  if (k0 != layers[0].K) return DN_ERR_UNSUPPORTED;
  for (int l = 0; l < n_layers; ++l) {
    const DnLayer& L = layers[l];
    const bool last = (l + 1 == n_layers);
    if (L.K % H_KS || L.K < H_KS || L.N % 32 || L.N < 32 || L.N > 256) return DN_ERR_UNSUPPORTED;
    if (L.emul || L.dots_src || L.head_w) return DN_ERR_UNSUPPORTED;
    if (L.relu_mask_src && (!last || L.residual || (reinterpret_cast<uintptr_t>(L.relu_mask_src) & 15)))
      return DN_ERR_UNSUPPORTED;
    if (L.bias && (reinterpret_cast<uintptr_t>(L.bias) & 15)) return DN_ERR_UNSUPPORTED;
    if (!last && (L.residual || L.row_scale)) return DN_ERR_UNSUPPORTED;
    if (!last && L.N % H_KS) return DN_ERR_UNSUPPORTED;
    if (L.residual && (L.res_scale != 1.f || L.ld_res % 4 || (reinterpret_cast<uintptr_t>(L.residual) & 15)))
      return DN_ERR_UNSUPPORTED;
    if (L.out && (L.ld_out % 4 || (reinterpret_cast<uintptr_t>(L.out) & 15))) return DN_ERR_UNSUPPORTED;
    if (l > 0 && L.K != layers[l - 1].N) return DN_ERR_UNSUPPORTED;
  }
  if (!layers[n_layers - 1].out) return DN_ERR_UNSUPPORTED;
  return DN_OK;
}

int tc_rows_chain16(const DnRowsSrc& src, const DnLayer* layers, int n_layers, int64_t V, int sm_count, cudaStream_t st) {
  if (V <= 0) return DN_OK;
  if (V >= (1ll << 31) - 256) return DN_ERR_UNSUPPORTED;
  int dev = 0;
  DN_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) return DN_ERR_UNSUPPORTED;
  if (g_attr_dev[dev] == 0) {
    g_attr_dev[dev] =
        cudaFuncSetAttribute(rows_chain16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM) == cudaSuccess ? 1 : -1;
    if (g_attr_dev[dev] < 0) cudaGetLastError();
  }
  if (g_attr_dev[dev] < 0) return DN_ERR_UNSUPPORTED;
  HParams p;
  HMaps maps;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.n_layers = n_layers;
  p.V = V;
  p.tile_group = layers[0].tile_group;
  p.group_stride = layers[0].group_stride;
  p.nsrc = src.nsrc;
  for (int s = 0; s < src.nsrc; ++s) {
    p.src_width[s] = src.width[s];
    if (h_box_map(&maps.src[s], src.ptr[s], src.width[s], src.ld[s], V, H_TILE)) return DN_ERR_UNSUPPORTED;
  }
  for (int l = 0; l < n_layers; ++l) {
    const DnLayer& L = layers[l];
    HLayer& T = p.layer[l];
    if (!L.prepacked || L.pack_fmt != 2) return DN_ERR_INVALID_ARGUMENT;
    T.wpack = L.prepacked; T.bias = L.bias; T.row_scale = L.row_scale; T.K = L.K; T.N = L.N; T.relu = L.relu;
    T.has_out = L.out != nullptr;
    T.has_res = L.residual ? 1 : (L.relu_mask_src ? 2 : 0);
    if (L.out && h_box_map(&maps.out[l], L.out, L.N, L.ld_out, V, 32)) return DN_ERR_UNSUPPORTED;
    if (L.residual && h_box_map(&maps.res, L.residual, L.N, L.ld_res, V, 32)) return DN_ERR_UNSUPPORTED;
    if (L.relu_mask_src && h_box_map(&maps.res, L.relu_mask_src, L.N, L.N, V, 32)) return DN_ERR_UNSUPPORTED;
  }
  const int64_t ntiles = (V + H_TILE - 1) / H_TILE;
  const int grid = (int)(ntiles < sm_count ? ntiles : sm_count);
  rows_chain16_kernel<<<grid, H_THREADS, H_SMEM, st>>>(p, maps);
  DN_LAUNCH_CHECK();
  return DN_OK;
}
