// Family 1: W8A8 GEMM for B200 -- FP8 (e4m3 x e4m3 -> f32) and INT8 (s8 x s8 -> s32) with FastDM's
// per-token (row) x per-channel (column) scaling, INT8 asymmetric zero-point correction, bias and an
// optional GELU, as one persistent warp-specialised tcgen05 kernel:
//
//   warp 0   : TMA producer  -- A [M,K] and B^T [N,K] tiles (128 B of K per stage row, 128B swizzle)
//              into a STAGES-deep shared-memory ring, mbarrier complete_tx signalling
//   warp 1   : MMA issuer    -- one elected thread issues tcgen05.mma (kind::f8f6f4 / kind::i8),
//              accumulators live in TMEM (2 x BN columns, double buffered across tiles);
//              tcgen05.commit releases smem stages and publishes finished accumulators
//   warps 2-9: epilogue      -- two warps per TMEM lane quarter, each taking half of the BN columns:
//              tcgen05.ld the 128 x BN accumulator (one row per thread), apply
//              sA[m]*sB[n] (+azp rank-1 correction) (+bias) (+GELU), round, 16-byte global stores.
//              Runs concurrently with the next tile's main loop.
//
// Numerics follow fastdm/kernel/torch/matrixmul.py (the parity oracle):
//   fp8 : out = T( (acc*sA)*sB + bias )                                     (:33, torch._scaled_mm)
//   int8: out = T( T( float(acc - azp*azp_adj) * (sA*sB) ) + bias )         (:67-74)
// Reference CUDA path being replaced: csrc/torch_bindings.cpp:24-160 -> csrc/gemm/*.cu (CUTLASS).
#include <stdlib.h>

#include <atomic>

#include "sm100.cuh"

namespace fdm {
using namespace sm100;

constexpr int kBM = 128;          // rows per tile (UMMA M, cta_group::1)
constexpr int kBK = 128;          // bytes (= 8-bit elements) of K per stage: one 128B swizzle row
constexpr int kUmmaK = 32;        // K per tcgen05.mma for 8-bit operands
constexpr int kGemmThreads = 320; // TMA warp + MMA warp + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kGroupM = 16;       // m-tiles per rasterisation group (L2 reuse of B)

// CG = 2: two CTAs of a cluster work on one 256-row tile as a pair (tcgen05.mma.cta_group::2, M = 256): each CTA stages
// its own 128 rows of A and only HALF of the B tile (BN / 2 weight rows); the tensor cores exchange the halves. Per
// k-block a CTA pulls 32 KB through L2 -> shared memory instead of 48 KB (128 x 256 tiles on every SM ask L2 for
// ~94 B/clk/SM, ~26 TB/s chip-wide at full MMA rate: the single-CTA kernel is L2-bandwidth-bound at ~79 % tensor-pipe
// activity), and the ring is 6 stages deep instead of 4.
template <int BN, int CG = 1>
struct GemmSmem {
  static constexpr int kStages = CG == 2 ? 6 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int kABytes = kBM * kBK;
  static constexpr int kBBytes = BN / CG * kBK;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRing = kStages * kStageBytes;
  // epilogue per-column parameters: scale_b, bias, azp_adj, gate
  static constexpr int kEpiOff = kRing;
  static constexpr int kEpiBytes = 4 * BN * 4 + BN * 2 + 16;  // ... + gate as bf16 + two "gate is not bf16" flags
  static constexpr int kBarOff = kEpiOff + kEpiBytes;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kTotal = kBarOff + kBarBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = 2 * BN;                    // 512 / 256 / 128
  static_assert(kTotal <= 232448, "gemm: shared-memory layout exceeds 227 KB");
};

struct GemmParams {
  const float* scale_a;
  const float* scale_b;
  const int32_t* azp_adj;
  const int32_t* azp;
  const void* bias;
  void* d;
  int M, N, K;
  int64_t ldd;
  int out_dtype;
  int act;
  int tiles_m, tiles_n;
  int vec_store;
  int vec32;  // rows of d (and of the residual) are 32-byte aligned: 256-bit loads / stores
  // fused residual epilogue: d = residual + gate[row / rows_per_batch, n] * T(linear)
  const float* gate;     // fp32 [batches, N] or NULL
  const void* residual;  // out_dtype [M, N] (row stride ldr) or NULL; may alias d
  int64_t ldr;
  int64_t rows_per_batch;
  int round_steps;       // 1: T(gate * x) before the residual add (bf16 tensor-op chain)
  int group_m;           // m-tiles per rasterisation group
};

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& tm, int& tn) {
  const int group_size = group_m * tiles_n;
  const int g = t / group_size;
  const int first_m = g * group_m;
  const int gm = min(group_m, tiles_m - first_m);
  const int r = t - g * group_size;
  tm = first_m + r % gm;
  tn = r / gm;
}

// GATED: the residual + gate * out epilogue is compiled in (its residual prefetch holds 64 registers; without it the
// epilogue double-buffers its TMEM reads instead)
template <bool INT8, int BN, int ACT, int CG, bool GATED>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_w8a8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using S = GemmSmem<BN, CG>;
  static_assert(CG == 1 || BN == 256, "the CTA-pair variant is built for 256-column tiles");
  // CTA pair: rank 0 (the leader) runs the MMA thread; the ring's full barriers and the accumulator-empty barriers live
  // in the leader's shared memory, stage-empty and accumulator-full barriers are signalled in both CTAs (multicast commit)
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  constexpr int kTileM = kBM * CG;   // rows per tile of the pair
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  float* s_sb = reinterpret_cast<float*>(smem + S::kEpiOff);
  float* s_bias = s_sb + BN;
  int32_t* s_adj = reinterpret_cast<int32_t*>(s_bias + BN);
  float* s_gate = reinterpret_cast<float*>(s_adj + BN);
  uint16_t* s_gate16 = reinterpret_cast<uint16_t*>(s_gate + BN);      // the same gate row as bf16 bit patterns
  int* s_gate_inexact = reinterpret_cast<int*>(s_gate16 + BN);        // [2], indexed by tile parity
  const uint32_t bar_base = base + S::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S::kStages + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + S::kBarOff + (2 * S::kStages + 4) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8 * CG);   // one arrival per epilogue warp of every CTA of the pair
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<CG>(smem_u32(tmem_ptr_smem), S::kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();  // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int first_tile = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        int tm, tn;
        tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
        const int m0 = tm * kTileM + (int)rank * kBM, n0 = tn * BN + (int)rank * (BN / CG);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * S::kStageBytes;
          const uint32_t sb = sa + S::kABytes;
          if (CG == 2) {
            // the leader's full barrier counts the bytes of both CTAs' boxes
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * S::kStageBytes);
            const uint32_t full_lead = mapa_u32(full_bar(stage), 0);
            tma_load_2d_cg2(sa, &tmap_a, full_lead, kb * kBK, m0);
            tma_load_2d_cg2(sb, &tmap_b, full_lead, kb * kBK, n0);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), S::kStageBytes);
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kBK, m0);
            tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kBK, n0);
          }
          if (++stage == S::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        int tm, tn;
        tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
        const int n0 = tn * BN;
        // (a pair always multiplies the full BN columns: each CTA contributes exactly BN / 2 weight rows, rows beyond N
        // are zero-filled by TMA and never stored)
        int n_eff = CG == 2 ? BN : min(BN, p.N - n0);
        n_eff = (n_eff + 15) & ~15;
        const uint32_t idesc =
            INT8 ? make_idesc(kFmtS8, kFmtS8, kAccS32, kTileM, (uint32_t)n_eff)
                 : make_idesc(kFmtE4M3, kFmtE4M3, kAccF32, kTileM, (uint32_t)n_eff);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * S::kStageBytes;
          const uint64_t adesc = make_desc_kmajor_sw128(sa);
          const uint64_t bdesc = make_desc_kmajor_sw128(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advancing K inside the 128B swizzle row: +32 bytes -> +2 in the (>>4) address field
            umma_ss<INT8 ? MmaKind::I8 : MmaKind::F8F6F4, CG>(d_tmem, adesc + 2u * k, bdesc + 2u * k,
                                                               idesc, (uint32_t)((kb | k) != 0));
          }
          if (CG == 2) tc_commit_cg2(empty_bar(stage), 0b11);  // the same stage in both CTAs
          else tc_commit(empty_bar(stage));
          if (++stage == S::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (CG == 2) tc_commit_cg2(tfull_bar(acc), 0b11);
        else tc_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int epi_tid = threadIdx.x - 64;
    const int lane_group = warp & 3;  // TMEM lanes [32*lane_group, +32) are this warp's
    const int col_half = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    const int row_in_tile = lane_group * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tile_par = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step, tile_par ^= 1) {
      int tm, tn;
      tile_coords(t, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
      const int m0 = tm * kTileM + (int)rank * kBM, n0 = tn * BN;   // this CTA's 128 rows of the tile
      if (epi_tid == 0) s_gate_inexact[tile_par] = 0;  // last read two tiles ago
      named_bar_sync(1, kEpiThreads);
      for (int i = epi_tid; i < BN; i += kEpiThreads) {
        const int col = n0 + i;
        const bool ok = col < p.N;
        s_sb[i] = ok ? p.scale_b[col] : 0.f;
        float b = 0.f;
        if (ok && p.bias != nullptr) {
          b = (p.out_dtype == FDM_BF16)
                  ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[col])
                  : __half2float(reinterpret_cast<const __half*>(p.bias)[col]);
        }
        s_bias[i] = b;
        if (INT8) s_adj[i] = (ok && p.azp_adj != nullptr) ? p.azp_adj[col] : 0;
        // the gate row of the tile's batch, staged once per tile: read per element from global it was two thirds
        // of the epilogue's load instructions (ncu: 313 k load requests against 98 k stores, lg_throttle stalls)
        if (GATED && p.gate != nullptr) {
          const float gv = ok ? p.gate[(int64_t)(m0 / p.rows_per_batch) * p.N + col] : 1.f;
          s_gate[i] = gv;
          s_gate16[i] = (uint16_t)(__float_as_uint(gv) >> 16);
          if ((__float_as_uint(gv) & 0xffffu) != 0u) s_gate_inexact[tile_par] = 1;
        }
      }
      named_bar_sync(1, kEpiThreads);

      const int row = m0 + row_in_tile;
      const bool row_ok = row < p.M;
      const float sa = row_ok ? p.scale_a[row] : 0.f;
      int zp = 0;
      if (INT8) zp = (row_ok && p.azp != nullptr) ? p.azp[row] : 0;
      const bool has_bias = p.bias != nullptr;

      // the residual rows do not depend on the accumulator: fetch this warp's share (BN/2 columns of its
      // row, 16-byte pieces) before waiting for the main loop, so their DRAM latency hides under it
      constexpr int kMyChunks = BN / 64;
      U128 resv[GATED ? kMyChunks : 1][4];
      if (GATED && p.residual != nullptr && row_ok) {
        const uint16_t* rrow = reinterpret_cast<const uint16_t*>(p.residual) + (int64_t)row * p.ldr + n0;
#pragma unroll
        for (int cc = 0; cc < kMyChunks; ++cc) {
          const int c = col_half * kMyChunks + cc;
          if (p.vec32 && n0 + c * 32 + 32 <= p.N) {
            ldg256(rrow + c * 32, resv[cc][0], resv[cc][1]);
            ldg256(rrow + c * 32 + 16, resv[cc][2], resv[cc][3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (n0 + c * 32 + q * 8 < p.N) resv[cc][q] = ldg128(rrow + c * 32 + q * 8);
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(acc * BN);
      const int n_valid = min(BN, p.N - n0);
      // the TMEM read of chunk c+1 is in flight while chunk c is processed (two register buffers): with two epilogue
      // warps per scheduler nothing else covers the tcgen05.ld latency (ncu: 25 % long-scoreboard stalls on it)
      constexpr bool kPipe = !GATED;
      uint32_t rbuf[kPipe ? 2 : 1][32];
      if (kPipe && col_half * kMyChunks * 32 < n_valid) tmem_ld_32x32(t_row + (uint32_t)(col_half * kMyChunks * 32), rbuf[0]);
#pragma unroll
      for (int cc = 0; cc < kMyChunks; ++cc) {
        const int c = col_half * kMyChunks + cc;
        if (c * 32 >= n_valid) break;
        if (!kPipe) tmem_ld_32x32(t_row + (uint32_t)(c * 32), rbuf[0]);
        tmem_ld_wait();
        if (kPipe && cc + 1 < kMyChunks && (c + 1) * 32 < n_valid)
          tmem_ld_32x32(t_row + (uint32_t)((c + 1) * 32), rbuf[kPipe ? (cc + 1) & 1 : 0]);
        uint32_t(&r)[32] = rbuf[kPipe ? cc & 1 : 0];
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sb = s_sb[c * 32 + j];
          const float bs = s_bias[c * 32 + j];
          float x;
          if (INT8) {
            const int a = (int)r[j] - zp * s_adj[c * 32 + j];
            x = (float)a * (sa * sb);
            if (has_bias) {
              // oracle rounds to the output dtype before the bias add (matrixmul.py:72-74)
              x = (p.out_dtype == FDM_BF16) ? round_to<__nv_bfloat16>(x) : round_to<__half>(x);
              x = x + bs;
            }
          } else {
            x = (__uint_as_float(r[j]) * sa) * sb + bs;
          }
          v[j] = x;
        }
        if (ACT != FDM_ACT_NONE) {
          // unfused reference: GEMM output is rounded to T, then F.gelu runs on that tensor
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = (p.out_dtype == FDM_BF16) ? round_to<__nv_bfloat16>(v[j]) : round_to<__half>(v[j]);
            v[j] = (ACT == FDM_ACT_GELU_TANH) ? gelu_tanh(x) : gelu_erf(x);
          }
        }
        if (row_ok) {
          const int col0 = n0 + c * 32;
          const bool bf = p.out_dtype == FDM_BF16;
          if (GATED && (p.gate != nullptr || p.residual != nullptr)) {
            // reference chains (flux.py:153-154,161-163,69-72; wan.py:97,105,112): the linear's output
            // is a T tensor, then gate * out (+ rounding for bf16 tensor ops), then residual + .
            // rows of one tile normally share a batch (and with it the staged gate row)
            const bool gate_staged = (m0 / p.rows_per_batch) == (min(m0 + kBM, p.M) - 1) / p.rows_per_batch;
            const float* grow = p.gate ? p.gate + (int64_t)(row / p.rows_per_batch) * p.N + col0 : nullptr;
            const bool has_res = p.residual != nullptr;
            // FLUX / Qwen / SD3.5 chain T(residual + T(gate * T(linear))) with a bf16-valued gate: three native
            // packed bf16 operations per pair of columns (see common.cuh bmul2), the residual stays packed
            const bool packed = bf && p.vec_store && p.round_steps && grow != nullptr && has_res && gate_staged &&
                                s_gate_inexact[tile_par] == 0 && col0 + 32 <= p.N;
            if (packed) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 g4 = *reinterpret_cast<const uint4*>(s_gate16 + c * 32 + q * 8);
                const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
                const uint32_t rw[4] = {resv[cc][q].x, resv[cc][q].y, resv[cc][q].z, resv[cc][q].w};
                uint32_t ow[4];
#pragma unroll
                for (int pr = 0; pr < 4; ++pr)
                  ow[pr] = badd2(rw[pr], bmul2(pack_bf16(v[q * 8 + 2 * pr], v[q * 8 + 2 * pr + 1]), gw[pr]));
                resv[cc][q] = U128{ow[0], ow[1], ow[2], ow[3]};
              }
              uint16_t* dst = reinterpret_cast<uint16_t*>(p.d) + (int64_t)row * p.ldd + col0;
              if (p.vec32) {
                stg256(dst, resv[cc][0], resv[cc][1]);
                stg256(dst + 16, resv[cc][2], resv[cc][3]);
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) stg128(dst + q * 8, resv[cc][q]);
              }
              continue;  // next 32-column chunk
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (col0 + q * 8 < p.N) {
                float gq[8], rq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  gq[j] = 1.f;
                  rq[j] = 0.f;
                }
                if (grow && gate_staged) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) gq[j] = s_gate[c * 32 + q * 8 + j];
                } else if (grow) {
                  const float4 g0 = *reinterpret_cast<const float4*>(grow + q * 8);
                  const float4 g1 = *reinterpret_cast<const float4*>(grow + q * 8 + 4);
                  gq[0] = g0.x; gq[1] = g0.y; gq[2] = g0.z; gq[3] = g0.w;
                  gq[4] = g1.x; gq[5] = g1.y; gq[6] = g1.z; gq[7] = g1.w;
                }
                if (has_res) {
                  if (bf) unpack8<__nv_bfloat16>(resv[cc][q], rq); else unpack8<__half>(resv[cc][q], rq);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float x = bf ? round_to<__nv_bfloat16>(v[q * 8 + j]) : round_to<__half>(v[q * 8 + j]);
                  if (grow) {
                    x = __fmul_rn(x, gq[j]);
                    if (p.round_steps) x = bf ? round_to<__nv_bfloat16>(x) : round_to<__half>(x);
                  }
                  v[q * 8 + j] = __fadd_rn(rq[j], x);
                }
              }
            }
          }
          if (p.vec32 && col0 + 32 <= p.N) {
            uint16_t* dst = reinterpret_cast<uint16_t*>(p.d) + (int64_t)row * p.ldd + col0;
            U128 o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (bf) {
                o[q].x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
                o[q].y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                o[q].z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
                o[q].w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
              } else {
                o[q].x = pack_f16(v[q * 8 + 0], v[q * 8 + 1]);
                o[q].y = pack_f16(v[q * 8 + 2], v[q * 8 + 3]);
                o[q].z = pack_f16(v[q * 8 + 4], v[q * 8 + 5]);
                o[q].w = pack_f16(v[q * 8 + 6], v[q * 8 + 7]);
              }
            }
            stg256(dst, o[0], o[1]);
            stg256(dst + 16, o[2], o[3]);
          } else if (p.vec_store) {
            uint16_t* dst = reinterpret_cast<uint16_t*>(p.d) + (int64_t)row * p.ldd + col0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (col0 + q * 8 < p.N) {
                U128 o;
                if (bf) {
                  o.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
                  o.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                  o.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
                  o.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
                } else {
                  o.x = pack_f16(v[q * 8 + 0], v[q * 8 + 1]);
                  o.y = pack_f16(v[q * 8 + 2], v[q * 8 + 3]);
                  o.z = pack_f16(v[q * 8 + 4], v[q * 8 + 5]);
                  o.w = pack_f16(v[q * 8 + 6], v[q * 8 + 7]);
                }
                stg128(dst + q * 8, o);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < p.N) {
                if (bf)
                  reinterpret_cast<__nv_bfloat16*>(p.d)[(int64_t)row * p.ldd + col0 + j] =
                      __float2bfloat16_rn(v[j]);
                else
                  reinterpret_cast<__half*>(p.d)[(int64_t)row * p.ldd + col0 + j] =
                      __float2half_rn(v[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && rank != 0) mbar_arrive_remote(tempty_bar(acc), 0);
        else mbar_arrive(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();  // neither CTA's shared memory / TMEM goes away while the pair is still working
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, S::kTmemCols);
  }
}

template <bool INT8, int BN, int ACT, int CG, bool GATED>
static int launch_gemm_g(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                         cudaStream_t st) {
  using S = GemmSmem<BN, CG>;
  static std::atomic<bool> attr_set[64];  // zero-initialised; setting the attribute twice is harmless
  int dev = 0;
  FDM_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    FDM_CUDA(cudaFuncSetAttribute(gemm_w8a8_kernel<INT8, BN, ACT, CG, GATED>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set[dev].store(true, std::memory_order_release);
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int slots = num_sms() / CG;   // CTAs, or CTA pairs
  const int grid = (tiles < slots ? tiles : slots) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FDM_CUDA(cudaLaunchKernelEx(&cfg, gemm_w8a8_kernel<INT8, BN, ACT, CG, GATED>, ta, tb, p));
  FDM_LAUNCH_CHECK("gemm_w8a8 kernel launch");
  return FDM_OK;
}

template <bool INT8, int BN, int ACT, int CG>
static int launch_gemm_a(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  if (p.gate != nullptr || p.residual != nullptr) return launch_gemm_g<INT8, BN, ACT, CG, true>(ta, tb, p, st);
  return launch_gemm_g<INT8, BN, ACT, CG, false>(ta, tb, p, st);
}

template <bool INT8, int BN, int CG = 1>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       cudaStream_t st) {
  if (p.act == FDM_ACT_GELU_TANH) return launch_gemm_a<INT8, BN, FDM_ACT_GELU_TANH, CG>(ta, tb, p, st);
  if (p.act == FDM_ACT_GELU_ERF) return launch_gemm_a<INT8, BN, FDM_ACT_GELU_ERF, CG>(ta, tb, p, st);
  return launch_gemm_a<INT8, BN, FDM_ACT_NONE, CG>(ta, tb, p, st);
}

static int gemm_common(bool int8, const void* a, const void* b, const float* scale_a,
                       const float* scale_b, const int32_t* azp_adj, const int32_t* azp,
                       const void* bias, void* d, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldd, int out_dtype, int act, const float* gate,
                       const void* residual, int64_t ldr, int64_t rows_per_batch, int round_steps,
                       void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  FDM_REQUIRE(M >= 0 && N >= 0 && K > 0, "gemm: bad shape M=%lld N=%lld K=%lld", (long long)M,
              (long long)N, (long long)K);
  if (M == 0 || N == 0) return FDM_OK;
  FDM_REQUIRE(a && b && scale_a && scale_b && d, "gemm: null pointer");
  FDM_REQUIRE(out_dtype == FDM_BF16 || out_dtype == FDM_F16, "gemm: out_dtype must be bf16 or f16");
  FDM_REQUIRE(act >= FDM_ACT_NONE && act <= FDM_ACT_GELU_ERF, "gemm: unknown activation %d", act);
  FDM_REQUIRE(K % 16 == 0, "gemm: K (%lld) must be a multiple of 16", (long long)K);
  FDM_REQUIRE(N % 8 == 0, "gemm: N (%lld) must be a multiple of 8", (long long)N);
  FDM_REQUIRE(lda >= K && ldb >= K && ldd >= N, "gemm: leading dimension too small");
  FDM_REQUIRE(lda % 16 == 0 && ldb % 16 == 0, "gemm: lda/ldb must be multiples of 16 bytes");
  FDM_REQUIRE((uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0,
              "gemm: a and b must be 16-byte aligned");
  FDM_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "gemm: dimension too large");
  FDM_REQUIRE((azp == nullptr) == (azp_adj == nullptr), "gemm: azp and azp_adj come together");

  // tile width: widest tile that still gives every SM work
  const int sms = num_sms();
  const int64_t tiles_m = (M + kBM - 1) / kBM;
  int bn = 256;
  if (tiles_m * ((N + 255) / 256) < sms) bn = 128;
  if (bn == 128 && tiles_m * ((N + 127) / 128) < sms) bn = 64;
  if (N <= 64) bn = 64;
  else if (N <= 128 && bn > 128) bn = 128;
  {  // experiments only (read once: thread-safe static initialisation, no getenv on the launch path)
    static const int forced_bn = [] { const char* e = getenv("FDM_GEMM_BN"); return e ? atoi(e) : 0; }();
    if (forced_bn == 64 || forced_bn == 128 || forced_bn == 256) bn = forced_bn;
  }
  // CTA pairs (256 x 256 tiles) once there are enough of them to give every SM pair a tile
  static const int pair_setting = [] { const char* e = getenv("FDM_GEMM_CG"); return e ? atoi(e) : 2; }();
  const int64_t tiles_m2 = (M + 2 * kBM - 1) / (2 * kBM);
  // (measured: with fewer tiles than pair slots -- M = 512, N = 9216: 72 on 74 -- pairs are no faster than 128 x 128 tiles)
  const int64_t tiles_n256 = (N + 255) / 256;
  const bool pair = pair_setting == 2 && N > 128 && tiles_m2 * tiles_n256 >= sms / 2;
  if (pair) bn = 256;
  const int64_t tiles_n = (N + bn - 1) / bn;
  const int64_t tiles_m_used = pair ? tiles_m2 : tiles_m;
  FDM_REQUIRE(tiles_m_used * tiles_n < (1LL << 31), "gemm: too many tiles");

  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda};
    uint32_t box[2] = {(uint32_t)kBK, (uint32_t)kBM};
    rc = make_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, a, dims, strides, box,
                   CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldb};
    uint32_t box[2] = {(uint32_t)kBK, (uint32_t)(pair ? bn / 2 : bn)};   // a CTA of a pair stages half of the weight rows
    rc = make_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b, dims, strides, box,
                   CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  GemmParams p;
  p.scale_a = scale_a;
  p.scale_b = scale_b;
  p.azp_adj = azp_adj;
  p.azp = azp;
  p.bias = bias;
  p.d = d;
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)K;
  p.ldd = ldd;
  p.out_dtype = out_dtype;
  p.act = act;
  p.tiles_m = (int)tiles_m_used;
  p.tiles_n = (int)tiles_n;
  p.vec_store = (ldd % 8 == 0 && (uintptr_t)d % 16 == 0) ? 1 : 0;
  p.vec32 = (ldd % 16 == 0 && (uintptr_t)d % 32 == 0 &&
             (residual == nullptr || (ldr % 16 == 0 && (uintptr_t)residual % 32 == 0))) ? 1 : 0;
  if (gate != nullptr || residual != nullptr) {
    FDM_REQUIRE(rows_per_batch > 0, "gemm: rows_per_batch must be positive");
    FDM_REQUIRE(gate == nullptr || (uintptr_t)gate % 16 == 0, "gemm: gate must be 16-byte aligned fp32");
    FDM_REQUIRE(residual == nullptr || (ldr >= N && ldr % 8 == 0 && (uintptr_t)residual % 16 == 0),
                "gemm: residual needs ldr >= N, ldr %% 8 == 0 and 16-byte alignment");
  }
  p.gate = gate;
  p.residual = residual;
  p.ldr = ldr;
  p.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  p.round_steps = round_steps;
  {
    static const int forced_gm = [] { const char* e = getenv("FDM_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();
    p.group_m = forced_gm > 0 ? forced_gm : kGroupM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (pair) return int8 ? launch_gemm<true, 256, 2>(ta, tb, p, st) : launch_gemm<false, 256, 2>(ta, tb, p, st);
  if (int8) {
    if (bn == 256) return launch_gemm<true, 256>(ta, tb, p, st);
    if (bn == 128) return launch_gemm<true, 128>(ta, tb, p, st);
    return launch_gemm<true, 64>(ta, tb, p, st);
  }
  if (bn == 256) return launch_gemm<false, 256>(ta, tb, p, st);
  if (bn == 128) return launch_gemm<false, 128>(ta, tb, p, st);
  return launch_gemm<false, 64>(ta, tb, p, st);
}

}  // namespace fdm

extern "C" {

int fdm_gemm_fp8(const void* a, const void* b, const float* scale_a, const float* scale_b,
                 const void* bias, void* d, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldd, int out_dtype, int act, void* stream) {
  return fdm::gemm_common(false, a, b, scale_a, scale_b, nullptr, nullptr, bias, d, M, N, K, lda,
                          ldb, ldd, out_dtype, act, nullptr, nullptr, 0, 1, 0, stream);
}

int fdm_gemm_int8(const void* a, const void* b, const float* scale_a, const float* scale_b,
                  const int32_t* azp_adj, const int32_t* azp, const void* bias, void* d, int64_t M,
                  int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd, int out_dtype,
                  int act, void* stream) {
  return fdm::gemm_common(true, a, b, scale_a, scale_b, azp_adj, azp, bias, d, M, N, K, lda, ldb,
                          ldd, out_dtype, act, nullptr, nullptr, 0, 1, 0, stream);
}

int fdm_gemm_fp8_residual(const void* a, const void* b, const float* scale_a, const float* scale_b,
                          const void* bias, void* d, int64_t M, int64_t N, int64_t K, int64_t lda,
                          int64_t ldb, int64_t ldd, int out_dtype, int act, const float* gate,
                          const void* residual, int64_t ldr, int64_t rows_per_batch, int round_steps,
                          void* stream) {
  return fdm::gemm_common(false, a, b, scale_a, scale_b, nullptr, nullptr, bias, d, M, N, K, lda,
                          ldb, ldd, out_dtype, act, gate, residual, ldr, rows_per_batch, round_steps,
                          stream);
}

int fdm_gemm_int8_residual(const void* a, const void* b, const float* scale_a, const float* scale_b,
                           const int32_t* azp_adj, const int32_t* azp, const void* bias, void* d,
                           int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd,
                           int out_dtype, int act, const float* gate, const void* residual,
                           int64_t ldr, int64_t rows_per_batch, int round_steps, void* stream) {
  return fdm::gemm_common(true, a, b, scale_a, scale_b, azp_adj, azp, bias, d, M, N, K, lda, ldb,
                          ldd, out_dtype, act, gate, residual, ldr, rows_per_batch, round_steps,
                          stream);
}

}  // extern "C"
