// Family 2: flash attention forward for B200 -- dense and block-sparse, non-causal MHA, bf16 / fp16
// operands, token-major ("NHD") layouts with arbitrary token / batch strides (q, k, v may be
// last-dim slices of a fused qkv projection).
//
// One CTA owns 256 query rows of one (batch, head): two 128-row Q tiles A and B that share every
// K / V tile pulled through shared memory (halves L2->SMEM traffic per flop and lets one tile's
// softmax overlap the other tile's MMAs):
//
//   warps 0-3 : softmax + correction + epilogue of Q tile A (thread <-> query row)
//   warps 4-7 : same for Q tile B
//   warp  8   : TMA producer -- Q tiles once, then the K / V tiles through an NS-stage ring
//               (128-byte swizzled [128 rows x 64 elements] boxes; OOB rows zero-filled)
//   warp  9   : MMA issuer (one elected thread runs the whole loop) -- S_X = Q_X K^T (SS, K-major x K-major) into TMEM,
//               then O_X += P_X V with V consumed MN-major straight from its row-major tile (no transpose anywhere).
//               Dense calls whose P does not alias S: warps 9, 10, 11 take turns group by group (see the issuer).
//
//   TMEM columns: S_A [0,128) | S_B [128,256) | O_A [256,256+HD) | O_B [384,384+HD)
//   P_X: over S_X in TMEM (TS MMA; fp8, masked or short hd-128 calls), in the spare TMEM columns next to O_X (hd 64),
//   or in shared memory (hd 128 CTA pairs: two CTAs of a cluster issue every MMA together as cta_group::2, M = 256).
//   In the last two cases S_X is free as soon as the softmax warps hold it in registers and QK_X(t+1) goes out ahead
//   of PV_X(t).
//
// Softmax is the usual online form in the exp2 domain with lazy rescaling: the running max only
// moves (and O is only rescaled through TMEM) when it grows by more than 2^8, so the correction
// is off the critical path for all but the first tiles.
//
// Block-sparse ("Sparge" / radial masks, fastdm/sparse/xsparse.py): an int8 mask
// [B, H, ceil(Sq/bq), ceil(Sk/bk)] (bq, bk in {64,128}); KV tiles whose mask entries are all zero
// for this CTA are skipped by all three roles (no TMA, no MMA, no softmax); partially masked tiles
// get -inf on the masked 64-column segments. Masked-out keys are excluded from the softmax.
//
// Semantics: fastdm/kernel/torch/attention.py:7-43 (F.scaled_dot_product_attention, non-causal),
// checked against the fp32 reference of tests/test_attention.py:23-63 at atol 1.8e-2 (:94).
// Replaces the library routes of fastdm/kernel/cuda/attention.py:149-261.
#include <stdlib.h>

#include <atomic>

#include "sm100.cuh"

namespace fdm {
using namespace sm100;

constexpr int kAttnThreads = 384;  // 2 softmax warpgroups + 1 warpgroup hosting the TMA and MMA warps
constexpr int kQTile = 128;   // rows per Q tile (UMMA M)
constexpr int kKvTile = 128;  // keys per KV tile (UMMA N of QK^T, K extent of PV)
constexpr int kMaxKvTiles = 8192;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
// The MMA issuer: one elected thread runs the whole issue loop (1) or the warp runs it converged and elects a lane per
// instruction (0). Inside `if (elect_one())` ptxas keeps descriptors, addresses and predicates on the uniform datapath:
// ~3 uniform instructions per tcgen05.mma instead of ~13 (ELECT, 2 VOTEU, 4 R2UR, UMOVs), which matters because the
// issuing warp shares its scheduler with two softmax warps (tools/softmax_lab.cu: +15-20 cycles per MMA under that load
// with the per-instruction election, +2-3 with one thread).
#ifndef FDM_ATTN_ONE_ISSUER
#define FDM_ATTN_ONE_ISSUER 1
#endif
constexpr bool kOneIssuer = FDM_ATTN_ONE_ISSUER != 0;
// dense calls whose P does not alias S: MMA groups are issued in order of readiness instead of round-robin (see the issuer)
#ifndef FDM_ATTN_DYN
#define FDM_ATTN_DYN 0
#endif
constexpr bool kDynIssue = FDM_ATTN_DYN != 0;
// dense calls: each P arrival releases the other Q tile's next QK before its own PV (see the issuer)
#ifndef FDM_ATTN_QK_FIRST
#define FDM_ATTN_QK_FIRST 0
#endif
constexpr bool kQkFirst = FDM_ATTN_QK_FIRST != 0;
// issue pacing window in batches of 2 MMAs (0 = off; see the issuer)
#ifndef FDM_ATTN_PACE
#define FDM_ATTN_PACE 0
#endif
constexpr int kPace = FDM_ATTN_PACE;
// dense calls whose P does not alias S: warps 9, 10, 11 take turns at issuing the MMA groups (see the issuer)
#ifndef FDM_ATTN_ROTATE
#define FDM_ATTN_ROTATE 1
#endif
constexpr bool kRotate = FDM_ATTN_ROTATE != 0 && kOneIssuer && kPace == 0;
constexpr int kIssuers = 3;

template <int HD, int ES, bool PS, int CG>
struct AttnSmem {
  static constexpr int kHalves = HD * ES / 128;         // 128-byte column groups per row
  static constexpr int kTileBytes = kKvTile * HD * ES;  // one Q / K / V tile
  static constexpr int kKvBytes = kTileBytes / CG;      // this CTA's share of a K or V tile (CTA pair: half)
  static constexpr int kPBytes = PS ? kQTile * kKvTile * ES : 0;  // one P tile staged in shared memory
  static constexpr int kStages = CG == 2 ? 6 : (PS ? (kTileBytes == 32768 ? 3 : 6) : (kTileBytes == 32768 ? 4 : 6));
  static constexpr int kQOff = 0;
  static constexpr int kPOff = 2 * kTileBytes;
  static constexpr int kKvOff = kPOff + 2 * kPBytes;
  static constexpr int kBarOff = kKvOff + kStages * kKvBytes;
  static constexpr int kNumBars = 1 + 2 * kStages + 8 + 4;   // q_full, ring full/empty, 8 pipeline barriers, 4 issue-pacing barriers
  static constexpr int kFlagsOff = kBarOff + kNumBars * 8 + 16;
  static constexpr int kTotal = kFlagsOff + kMaxKvTiles / 8 + 1024;  // one flag bit per KV tile; 1 KB alignment slack
  static_assert(kTotal <= 232448, "attention: shared-memory layout exceeds 227 KB");
};

struct AttnParams {
  void* o;
  int64_t o_bs, o_ts;
  int o_vec32;
  const int8_t* mask;
  int B, H, Sq, Sk;
  int n_kv_tiles;
  int mask_bq, mask_bk, nbq, nbk;
  float scale_log2;
  int stagger;       // cycles by which Q tile B's first QK is held back (readiness-driven issue order)
  long long* trace;  // debug: per-event clock64 stamps of CTA (0,0,0), or nullptr
};

// Extra kernel parameter of the SCATTER kernels (Ulysses): the epilogue writes each query row straight into the output
// buffer of the rank that owns the row's token shard -- peer memory over NVLink -- instead of a local buffer that an
// all-to-all then redistributes: row r goes to o_peer[r / rows_per_peer], row r % rows_per_peer, token stride o_ts.
// A parameter of its own (and a kernel entry of its own) so that the ordinary kernels' parameter space, and with it
// their generated code, stay exactly as measured.
struct ScatterParams {
  void* o_peer[8];
  int rows_per_peer;
};

// debug timeline: trace[(role * 8 + event) * 64 + tile] = clock64(); role 0/1 = softmax warp 0 of
// Q tile A/B, role 2 (and 3) = MMA thread(s). Only CTA (0,0,0) writes, only when a buffer was registered.
constexpr int kTraceTiles = 64;
// (the stamps are compiled into a separate TRACE instantiation of the kernel: even predicated off they cost
// the production kernel 5-15 %)
template <bool TRACE>
__device__ __forceinline__ void trace_ev_t(const AttnParams& p, bool on, int role, int ev, uint32_t tile) {
  if constexpr (TRACE) {
    if (on && tile < kTraceTiles) p.trace[(role * 8 + ev) * kTraceTiles + tile] = clock64();
  }
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// operand dtype of q/k/v (and of P): 0 = bf16, 1 = fp16, 2 = fp8 e4m3
constexpr int kDtBF16 = 0, kDtF16 = 1, kDtE4M3 = 2;
template <int DT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return DT == kDtF16 ? pack_f16(lo, hi) : pack_bf16(lo, hi);
}

// per-warpgroup register re-budgeting: the kernel is launched at 168 regs/thread (64512 in all), so
// 256 x 208 + 128 x R must not exceed that or setmaxnreg.inc never completes: R <= 88
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ---- packed 2 x fp32 arithmetic (sm_100: FFMA2 / FADD2) and 3-input max -------------------------
__device__ __forceinline__ uint64_t f2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unf2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// max over 32 fp32 values held as raw bits, as a shallow tree (one softmax warp per scheduler is
// active at a time, so instruction-level parallelism is all the latency hiding there is)
__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    m[i] = fmax3(__uint_as_float(r[3 * i]), __uint_as_float(r[3 * i + 1]), __uint_as_float(r[3 * i + 2]));
  m[10] = fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31]));
  const float a = fmax3(m[0], m[1], m[2]), b = fmax3(m[3], m[4], m[5]), c = fmax3(m[6], m[7], m[8]);
  return fmax3(fmax3(a, b, c), m[9], m[10]);
}
// 2^x for a pair on the FMA pipe instead of MUFU: round-to-nearest split x = n + f, f in [-.5,.5],
// degree-3 minimax polynomial for 2^f (max rel. error 7.5e-5, far below the bf16 rounding of P),
// exponent patched in with an integer add. Takes load off the 16-op/clk/SM MUFU unit, which is
// otherwise exactly as busy as the tensor cores for hd = 128 attention.
__device__ __forceinline__ void ex2_emulated_pair(uint64_t y, float& p0, float& p1) {
  float y0, y1;
  unf2(y, y0, y1);
  y = f2(fmaxf(y0, -126.f), fmaxf(y1, -126.f));
  const uint64_t kMagic = f2(12582912.f, 12582912.f);      // 1.5 * 2^23
  const uint64_t kNegMagic = f2(-12582912.f, -12582912.f);
  const uint64_t kNegOne = f2(-1.f, -1.f);
  const uint64_t t = fadd2(y, kMagic);                      // round(y) in the low mantissa bits
  const uint64_t n = fadd2(t, kNegMagic);
  const uint64_t f = ffma2(n, kNegOne, y);                  // y - round(y)
  uint64_t q = ffma2(f, f2(0.05517186224460602f, 0.05517186224460602f), f2(0.2426111400127411f, 0.2426111400127411f));
  q = ffma2(q, f, f2(0.6932609677314758f, 0.6932609677314758f));
  q = ffma2(q, f, f2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  unf2(q, q0, q1);
  unf2(t, t0, t1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

template <int HD, int DT, int EMUX, bool PS, int CG, bool MASKED, bool TRACE, bool SCATTER>
__device__ __forceinline__ void attn_fwd_body(const CUtensorMap& tmap_q, const CUtensorMap& tmap_k, const CUtensorMap& tmap_v,
                                              const AttnParams& p, const ScatterParams& sp) {
  constexpr int ES = DT == kDtE4M3 ? 1 : 2;  // operand element size
  constexpr bool F16 = DT == kDtF16;
  constexpr int EMU = EMUX & 63;           // exponentials per 32 that run on the FMA pipe
  constexpr bool LS = (EMUX & 64) != 0;    // late store: the whole P tile is packed in registers before it is written
  static_assert(CG == 1 || (PS && HD == 128 && ES == 2), "the CTA-pair variant is built for hd 128, 16-bit operands, P in smem");
  using S = AttnSmem<HD, ES, PS, CG>;
  // hd 64 leaves half of each O block of TMEM unused: P goes there (PT) instead of over S, so -- exactly as with
  // P in shared memory (PS) -- S_X is free as soon as the softmax warps hold it in registers and QK_X(t+1) is
  // issued ahead of PV_X(t) (DEC): no smem needed, P still feeds a TS MMA.
  constexpr bool PT = !PS && HD == 64 && ES == 2;
  constexpr bool DEC = PS || PT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bar_base = base + S::kBarOff;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + S::kStages + s); };
  auto s_full = [&](int x) { return bar_base + 8u * (1 + 2 * S::kStages + x); };
  auto p_ready = [&](int x) { return bar_base + 8u * (1 + 2 * S::kStages + 2 + x); };
  auto o_done = [&](int x) { return bar_base + 8u * (1 + 2 * S::kStages + 4 + x); };
  auto s_free = [&](int x) { return bar_base + 8u * (1 + 2 * S::kStages + 6 + x); };
  auto pace_bar = [&](uint32_t b) { return bar_base + 8u * (1 + 2 * S::kStages + 8 + (b & 3u)); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + S::kBarOff + S::kNumBars * 8);
  uint8_t* flags = smem + S::kFlagsOff;

  // CTA pair (CG == 2): the two CTAs of a cluster own 256 query rows each and issue every MMA together as one
  // M = 256 tcgen05.mma.cta_group::2 -- each CTA stages only half of every K tile (64 keys) and half of
  // every V tile (64 head-dim columns), so the K/V ring costs half the shared memory and half the L2 reads.
  // Rank 0 (the leader) runs the MMA thread; kv_full / s_free / p_ready live in the leader's shared memory.
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  auto lead = [&](uint32_t bar) { return CG == 2 ? mapa_u32(bar, 0) : bar; };
  auto arrive_lead = [&](uint32_t bar) {  // arrive on the leader CTA's copy of a barrier
    if (rank == 0) mbar_arrive(bar);
    else mbar_arrive_remote(bar, 0);
  };
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * kQTile;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  constexpr bool has_mask = MASKED;  // a block mask was passed (p.mask != nullptr)
  auto trace_ev = [&](const AttnParams& pp, bool on, int role, int ev, uint32_t tile) { trace_ev_t<TRACE>(pp, on, role, ev, tile); };
  const int8_t* mask_bh = has_mask ? p.mask + ((int64_t)b * p.H + h) * p.nbq * p.nbk : nullptr;

  // TRACE: CTA life-cycle stamps in trace[(3 * 8 + 7) * 64 + i]: 0 kernel entry, 1 set-up done, 2 epilogue start, 3 epilogue
  // end, 4 kernel exit
  const bool life = TRACE && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
  if (life) p.trace[(3 * 8 + 7) * kTraceTiles + 0] = clock64();
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (uint32_t i = 0; i < 4; ++i) mbar_init(pace_bar(i), 1);
    for (int x = 0; x < 2; ++x) {
      mbar_init(s_full(x), 1);
      mbar_init(o_done(x), 1);
      // single CTA: one arrival per softmax thread; CTA pair: one elected arrival per softmax warp of both CTAs
      mbar_init(p_ready(x), CG == 2 ? 8 : 128);
      mbar_init(s_free(x), CG == 2 ? 8 : 128);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<CG>(smem_u32(tmem_ptr_smem), 512);
  if (has_mask) {
    // which KV tiles does this CTA (pair) need at all? (identical answer for all roles; one bit per tile)
    const int q0g = (int)(blockIdx.x / CG * CG) * 2 * kQTile;
    const int qb_lo = q0g / p.mask_bq;
    const int qb_hi = max(qb_lo, min((min(q0g + 2 * CG * kQTile, p.Sq) - 1) / p.mask_bq, p.nbq - 1));
    // (whole 32-bit words are written, zero beyond the last tile: next_active scans them with ffs)
    for (int j8 = threadIdx.x; j8 < (p.n_kv_tiles + 31) / 32 * 4; j8 += kAttnThreads) {
      uint32_t bits = 0;
      for (int jj = 0; jj < 8 && j8 * 8 + jj < p.n_kv_tiles; ++jj) {
        const int j = j8 * 8 + jj;
        const int kb_lo = (j * kKvTile) / p.mask_bk;
        const int kb_hi = min((min((j + 1) * kKvTile, p.Sk) - 1) / p.mask_bk, p.nbk - 1);
        int any = 0;
        for (int qb = qb_lo; qb <= qb_hi; ++qb)
          for (int kb = kb_lo; kb <= kb_hi; ++kb) any |= mask_bh[(int64_t)qb * p.nbk + kb];
        bits |= any ? (1u << jj) : 0u;
      }
      flags[j8] = (uint8_t)bits;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();  // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (life) p.trace[(3 * 8 + 7) * kTraceTiles + 1] = clock64();
  auto tile_active = [&](int j) -> bool { return !has_mask || ((flags[j >> 3] >> (j & 7)) & 1) != 0; };
  auto next_active = [&](int j) {  // first active tile >= j, or n_kv_tiles
    if (!has_mask) return j;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(flags);
    while (j < p.n_kv_tiles) {
      const uint32_t w = words[j >> 5] >> (j & 31);
      if (w != 0u) return j + __ffs((int)w) - 1;
      j = (j | 31) + 1;
    }
    return p.n_kv_tiles;
  };

  constexpr uint32_t kFmt = DT == kDtE4M3 ? kFmtE4M3 : (F16 ? kFmtF16 : kFmtBF16);
  constexpr MmaKind kKind = DT == kDtE4M3 ? MmaKind::F8F6F4 : MmaKind::F16;
  constexpr uint32_t kIdescQK = make_idesc(kFmt, kFmt, kAccF32, kQTile * CG, kKvTile, 0, 0);
  constexpr uint32_t kIdescPV = make_idesc(kFmt, kFmt, kAccF32, kQTile * CG, HD, 0, 1);
  constexpr int kBoxElems = 128 / ES;   // elements per 128-byte swizzled row
  constexpr int kKeysPerPV = 32 / ES;   // keys per P.V tcgen05.mma (K = 32 bytes of operand)

  if (warp >= 8) {
    reg_dealloc<88>();
  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * CG * S::kTileBytes);
      const uint32_t q_full_lead = lead(q_full);
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int hf = 0; hf < S::kHalves; ++hf) {
          const uint32_t dst = base + S::kQOff + x * S::kTileBytes + hf * (kQTile * 128);
          if (CG == 2) tma_load_4d_cg2(dst, &tmap_q, q_full_lead, hf * kBoxElems, h, q0 + x * kQTile, b);
          else tma_load_4d(dst, &tmap_q, q_full, hf * kBoxElems, h, q0 + x * kQTile, b);
        }
      uint32_t u = 0;
      auto load_tile = [&](const CUtensorMap* tm, int j) {
        const int stage = u % S::kStages;
        const uint32_t parity = ((u / S::kStages) & 1u) ^ 1u;
        mbar_wait(kv_empty(stage), parity);
        // the full barrier (the leader's, for a CTA pair) counts the bytes of the whole tile
        if (rank == 0) mbar_arrive_expect_tx(kv_full(stage), S::kTileBytes);
        const uint32_t dst = base + S::kKvOff + stage * S::kKvBytes;
        if (CG == 2) {
          const uint32_t full_lead = lead(kv_full(stage));
          if (tm == &tmap_k) {
            // this CTA's 64 keys of the tile, both 64-element head-dim panels ([64 keys x 128 B] each)
#pragma unroll
            for (int hf = 0; hf < S::kHalves; ++hf)
              tma_load_4d_cg2(dst + hf * (kKvTile / 2 * 128), tm, full_lead, hf * kBoxElems, h,
                              j * kKvTile + (int)rank * (kKvTile / 2), b);
          } else {
            // all 128 keys, this CTA's 64 head-dim columns (one [128 keys x 128 B] panel)
            tma_load_4d_cg2(dst, tm, full_lead, (int)rank * kBoxElems, h, j * kKvTile, b);
          }
        } else {
#pragma unroll
          for (int hf = 0; hf < S::kHalves; ++hf)
            tma_load_4d(dst + hf * (kKvTile * 128), tm, kv_full(stage), hf * kBoxElems, h, j * kKvTile, b);
        }
        ++u;
      };
      if (DEC) {
        // ring order = order of first use by the MMA warp: K(0), then K(t+1), V(t) for every active tile t
        int j = next_active(0);
        if (j < p.n_kv_tiles) load_tile(&tmap_k, j);
        while (j < p.n_kv_tiles) {
          const int jn = next_active(j + 1);
          if (jn < p.n_kv_tiles) load_tile(&tmap_k, jn);
          load_tile(&tmap_v, j);
          j = jn;
        }
      } else {
        for (int j = next_active(0); j < p.n_kv_tiles; j = next_active(j + 1)) {
          load_tile(&tmap_k, j);
          load_tile(&tmap_v, j);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9 || (kRotate && DEC && !MASKED && !TRACE)) {
    // ===================== MMA issuer (the leader CTA's, for a pair) =====================
    // One thread issues everything: the tensor pipe runs one thread's MMAs back to back at 64 cycles each,
    // but drops to ~87 cycles when two threads' MMAs interleave (tools/cg2_rate.cu). A tcgen05.mma blocks
    // its thread once ~4 are queued (tools/mma_queue.cu), so whatever the thread does between two groups
    // of MMAs has to fit under ~256 cycles of queued work or the pipe drains.
    // kOneIssuer: one elected thread runs the loop; else the whole warp runs it converged and the MMAs and commits
    // themselves go out from one elected lane.
    if (rank == 0 && (!kOneIssuer || elect_one())) {
      // per-Q-tile operands as scalars (x is a run-time value for the two PS issuers)
      auto tS_of = [&](int x) { return tmem_base + (uint32_t)x * 128u; };
      auto tO_of = [&](int x) { return tmem_base + 256u + (uint32_t)x * 128u; };
      // descriptors: the start-address field is the low 14 bits (>>4), so stepping inside a tile is an add
      auto q_desc_of = [&](int x) { return make_desc_kmajor_sw128(base + S::kQOff + (uint32_t)x * S::kTileBytes); };
      auto p_desc_of = [&](int x) { return make_desc_kmajor_sw128(base + S::kPOff + (uint32_t)x * S::kPBytes); };
      auto commit = [&](uint32_t bar) {
        if (kOneIssuer) {
          if (CG == 2) tc_commit_cg2(bar, 0b11);  // the same barrier in both CTAs of the pair
          else tc_commit(bar);
        } else {
          if (CG == 2) tc_commit_cg2_elect(bar, 0b11);
          else tc_commit_elect(bar);
        }
      };
      // Issue pacing: a tcgen05.mma whose thread already has ~4 queued blocks at issue, and while it is blocked the softmax
      // warps on the same scheduler make almost no progress (per-warp arrival stamps: the two warps that share warp 9's
      // scheduler arrive 400-1600 cycles after the other six, and every P hand-over waits for the last warp). So the
      // thread never lets more than kPace batches of 2 MMAs be outstanding: each batch commits to one of four pacing
      // barriers and batch b first waits (mbarrier.try_wait: the thread is suspended, not blocked at issue) for batch
      // b - kPace to have completed.
      uint32_t pace_b = 0;
      auto pace_before = [&]() {
        if constexpr (kPace > 0 && kOneIssuer) {
          if (pace_b >= (uint32_t)kPace) {
            const uint32_t w = pace_b - (uint32_t)kPace;
            mbar_wait(pace_bar(w), (w >> 2) & 1u);
          }
        }
      };
      auto pace_after = [&]() {
        if constexpr (kPace > 0 && kOneIssuer) {
          if (CG == 2) tc_commit_cg2(pace_bar(pace_b), 0b01);  // (this CTA's barrier only: the issuing thread is the one that waits)
          else tc_commit(pace_bar(pace_b));
          ++pace_b;
        }
      };
      // `probe` runs after the 7th MMA of a group -- with the queue full, i.e. for free (one late probe measured
      // no worse than two earlier ones and succeeds more often)
      auto issue_qk = [&](int x, uint32_t k_smem, auto&& probe) {
        const uint64_t k_desc = make_desc_kmajor_sw128(k_smem);
        const uint64_t q_desc = q_desc_of(x);
        const uint32_t tS = tS_of(x);
#pragma unroll
        for (int ks = 0; ks < HD * ES / 32; ++ks) {  // 32 bytes of head dim per MMA
          const uint64_t off = (uint64_t)(((ks / 4) * (kQTile * 128) + (ks % 4) * 32) >> 4);
          // a CTA of a pair holds kKvTile / 2 keys per head-dim panel
          const uint64_t koff = (uint64_t)(((ks / 4) * (kKvTile / CG * 128) + (ks % 4) * 32) >> 4);
          if ((ks & 1) == 0) pace_before();
          umma_ss<kKind, CG, !kOneIssuer>(tS, q_desc + off, k_desc + koff, kIdescQK, ks != 0);
          if ((ks & 1) == 1) pace_after();
          if (ks == 6) probe();
        }
      };
      auto no_probe = [] {};
      auto issue_pv = [&](int x, uint32_t v_smem, bool accumulate, auto&& probe) {
        const uint64_t v_desc = make_desc_mnmajor_sw128(v_smem, kKvTile * 128, 1024);
        const uint64_t p_desc = p_desc_of(x);
        const uint32_t tO = tO_of(x);
        const uint32_t tP = PT ? tO + 64u : tS_of(x);  // PT: columns [320,384) / [448,512), else over S_X
#pragma unroll
        for (int ks = 0; ks < kKvTile / kKeysPerPV; ++ks) {
          // kKeysPerPV keys = that many 128-byte rows of V; the matching slice of P is 32 bytes of its
          // K-major rows in shared memory (PS) or 8 TMEM columns
          const uint64_t voff = (uint64_t)((ks * kKeysPerPV * 128) >> 4);
          const uint32_t acc = (accumulate || ks != 0) ? 1u : 0u;
          if ((ks & 1) == 0) pace_before();
          if (PS) {
            const uint64_t poff = (uint64_t)(((ks / 4) * (kQTile * 128) + (ks % 4) * 32) >> 4);
            umma_ss<kKind, CG, !kOneIssuer>(tO, p_desc + poff, v_desc + voff, kIdescPV, acc);
          } else {
            umma_ts<kKind, !kOneIssuer>(tO, tP + (uint32_t)ks * 8u, v_desc + voff, kIdescPV, acc);
          }
          if ((ks & 1) == 1) pace_after();
          if (ks == (kKvTile / kKeysPerPV) * 7 / 8 - 1) probe();
        }
      };
      auto stage_addr = [&](uint32_t u) { return base + S::kKvOff + (u % S::kStages) * S::kKvBytes; };
      auto wait_full = [&](uint32_t u) {
        mbar_wait(kv_full(u % S::kStages), (u / S::kStages) & 1u);
        tc_fence_after();
      };
      const bool tr = TRACE && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (kOneIssuer || lane == 0);

      int j = next_active(0);
      if (j < p.n_kv_tiles) {
        mbar_wait(q_full, 0);
        tc_fence_after();
        wait_full(0);
        if (DEC && !MASKED && !TRACE && kRotate) {
          // Rotating issuers (dense calls). A tcgen05.mma whose thread already has ~4 queued blocks at issue, and while it is
          // blocked the two softmax warps on the same scheduler make almost no progress (per-warp arrival stamps with one
          // issuer on warp 9: warps 1 and 5 hand their P rows over 400-1600 cycles after the other six, and every PV waits
          // for the last row). Pacing the issue through commit barriers starves the pipe (-10 %), so the stall is spread
          // instead: warps 9, 10 and 11 -- one per scheduler next to the TMA warp's -- run the same group sequence
          //     QK_A(0) QK_B(0) | QK_A(t+1) PV_A(t) QK_B(t+1) PV_B(t) | ...
          // and group g is issued by issuer g mod 3, after a token from the issuer of group g - 1 (issue order = pipe order;
          // a commit only tracks its own thread's MMAs, which is enough because the pipe completes them in issue order).
          // Every issuer waits itself for the barriers of the groups it issues.
          const uint32_t n = (uint32_t)p.n_kv_tiles;
          const uint32_t me = (uint32_t)(warp - 9);
          auto item_k = [&](uint32_t t) { return t == 0 ? 0u : 2u * t - 1u; };
          auto item_v = [&](uint32_t t) { return t + 1 < n ? 2u * t + 2u : 2u * n - 1u; };
          auto item_stage = [&](uint32_t it) { return it % (uint32_t)S::kStages; };
          auto item_par = [&](uint32_t it) { return (it / (uint32_t)S::kStages) & 1u; };
          auto stage_smem = [&](uint32_t st) { return base + S::kKvOff + st * S::kKvBytes; };
          uint32_t g = 0, mine = 0;   // group counter, number of groups this issuer has taken
          auto my_turn = [&]() {      // is group g mine? if so, wait for the token of group g - 1
            const bool yes = g % (uint32_t)kIssuers == me;
            if (yes && g > 0) mbar_wait(pace_bar(me), mine & 1u);
            return yes;
          };
          auto pass_on = [&]() {      // group g is in the queue: the next issuer may go
            mbar_arrive(pace_bar((me + 1u) % (uint32_t)kIssuers));
            ++mine;
          };
          // group 0 goes out without a token, so issuer 0's token count is one behind the others'
          if (my_turn()) {
            issue_qk(0, stage_smem(0), no_probe);
            commit(s_full(0));
            pass_on();
            if (me == 0) --mine;
          }
          ++g;
          if (my_turn()) {
            issue_qk(1, stage_smem(0), no_probe);
            commit(s_full(1));
            commit(kv_empty(0));
            pass_on();
          }
          ++g;
          for (uint32_t t = 0; t < n; ++t) {
            const uint32_t ph = t & 1u;
            const uint32_t iv = item_v(t), ik = item_k(t + 1);
            const bool has_next = t + 1 < n;
            if (has_next) {
              if (my_turn()) {   // QK_A(t+1)
                mbar_wait(kv_full(item_stage(ik)), item_par(ik));
                mbar_wait(s_free(0), ph);
                tc_fence_after();
                issue_qk(0, stage_smem(item_stage(ik)), no_probe);
                commit(s_full(0));
                pass_on();
              }
              ++g;
            }
            if (my_turn()) {     // PV_A(t)
              mbar_wait(kv_full(item_stage(iv)), item_par(iv));
              mbar_wait(p_ready(0), ph);
              tc_fence_after();
              issue_pv(0, stage_smem(item_stage(iv)), t != 0, no_probe);
              commit(o_done(0));
              pass_on();
            }
            ++g;
            if (has_next) {
              if (my_turn()) {   // QK_B(t+1)
                mbar_wait(kv_full(item_stage(ik)), item_par(ik));
                mbar_wait(s_free(1), ph);
                tc_fence_after();
                issue_qk(1, stage_smem(item_stage(ik)), no_probe);
                commit(s_full(1));
                commit(kv_empty(item_stage(ik)));
                pass_on();
              }
              ++g;
            }
            if (my_turn()) {     // PV_B(t)
              mbar_wait(kv_full(item_stage(iv)), item_par(iv));
              mbar_wait(p_ready(1), ph);
              tc_fence_after();
              issue_pv(1, stage_smem(item_stage(iv)), t != 0, no_probe);
              commit(o_done(1));
              commit(kv_empty(item_stage(iv)));
              pass_on();
            }
            ++g;
          }
        } else if (DEC && !MASKED && kQkFirst) {
          // Dense calls, QK-first order. What the softmax warps of Q tile X wait for at the end of tile t is S_X(t+1); what
          // they need ~1800 cycles later is their P buffer back (PV_X(t) done). With the round-robin order below, the thread
          // that has just seen P_B(t-1) issues PV_B(t-1) and only then QK_A(t+1): S_A(t+1) queues behind 8 MMAs it does not
          // depend on and arrives ~300-500 cycles after tile A's softmax wanted it (in-kernel timeline: that wait, per tile,
          // on both Q tiles). Here each P arrival releases the OTHER tile's next QK first and the PV second:
          //     wait P_A(t) | QK_B(t+1) | PV_A(t) | wait P_B(t) | QK_A(t+2) | PV_B(t)
          // The blocking waits keep the two softmax warpgroups half a tile apart (B's next S is only issued once A has
          // finished a tile and vice versa), which matters: in phase, both hit the MUFU unit at once and a tile's
          // exponentials take 1950 cycles instead of 1300 (measured with the readiness-driven order).
          // Ring items as above: K(0) | K(t+1), V(t) | ...; K(t+2) is first used one iteration before QK_B(t+2) frees it.
          const uint32_t n = (uint32_t)p.n_kv_tiles;
          auto item_k = [&](uint32_t t) { return t == 0 ? 0u : 2u * t - 1u; };
          auto item_v = [&](uint32_t t) { return t + 1 < n ? 2u * t + 2u : 2u * n - 1u; };
          auto item_stage = [&](uint32_t it) { return it % (uint32_t)S::kStages; };
          auto item_par = [&](uint32_t it) { return (it / (uint32_t)S::kStages) & 1u; };
          auto stage_smem = [&](uint32_t st) { return base + S::kKvOff + st * S::kKvBytes; };
          issue_qk(0, stage_smem(0), no_probe);
          commit(s_full(0));
          issue_qk(1, stage_smem(0), no_probe);
          commit(s_full(1));
          commit(kv_empty(0));
          if (n > 1) {
            const uint32_t ik = item_k(1);
            mbar_wait(kv_full(item_stage(ik)), item_par(ik));
            mbar_wait(s_free(0), 0u);
            tc_fence_after();
            issue_qk(0, stage_smem(item_stage(ik)), no_probe);
            commit(s_full(0));
          }
          for (uint32_t t = 0; t < n; ++t) {
            const uint32_t ph = t & 1u;
            const uint32_t iv = item_v(t);
            // ---- P_A(t) written: QK_B(t+1), then PV_A(t) ----
            trace_ev(p, tr, 2, 4, t);
            mbar_wait(p_ready(0), ph);
            trace_ev(p, tr, 2, 0, t);
            if (t + 1 < n) {
              const uint32_t ik = item_k(t + 1);
              mbar_wait(s_free(1), ph);   // S_B(t) is in registers (K(t+1) landed before QK_A(t+1) went out)
              tc_fence_after();
              issue_qk(1, stage_smem(item_stage(ik)), no_probe);
              commit(s_full(1));
              commit(kv_empty(item_stage(ik)));
            }
            trace_ev(p, tr, 2, 1, t);
            mbar_wait(kv_full(item_stage(iv)), item_par(iv));
            tc_fence_after();
            trace_ev(p, tr, 2, 2, t);
            issue_pv(0, stage_smem(item_stage(iv)), t != 0, no_probe);
            commit(o_done(0));
            trace_ev(p, tr, 2, 3, t);
            // ---- P_B(t) written: QK_A(t+2), then PV_B(t) ----
            mbar_wait(p_ready(1), ph);
            trace_ev(p, tr, 3, 0, t);
            if (t + 2 < n) {
              const uint32_t ik = item_k(t + 2);
              mbar_wait(kv_full(item_stage(ik)), item_par(ik));
              mbar_wait(s_free(0), ph ^ 1u);   // S_A(t+1) is in registers
              tc_fence_after();
              issue_qk(0, stage_smem(item_stage(ik)), no_probe);
              commit(s_full(0));
            }
            trace_ev(p, tr, 3, 1, t);
            tc_fence_after();
            trace_ev(p, tr, 3, 2, t);
            issue_pv(1, stage_smem(item_stage(iv)), t != 0, no_probe);
            commit(o_done(1));
            commit(kv_empty(item_stage(iv)));
            trace_ev(p, tr, 3, 3, t);
          }
        } else if (DEC && !MASKED && kDynIssue) {
          // Readiness-driven issue order (dense calls): the four groups of a tile -- QK_A(t+1), QK_B(t+1), PV_A(t), PV_B(t) --
          // have independent conditions (S_X free + K landed / P_X written + V landed), and with a fixed round-robin
          // order a group whose condition was met long ago waits behind one that is still blocked: QK_X(t+1) went out
          // ~1800 cycles after S_X became free, the softmax warps then waited ~500 cycles per tile for S and ~300 for
          // their P buffer (in-kernel timeline). Here the thread polls all four conditions and issues whichever group
          // is ready; the softmax of tile t then always finds S(t+1) and a free P buffer. Issue itself is cheap
          // (one thread, uniform datapath), so the polling does not hold the tensor pipe back.
          // Ring items (dense): K(0) | K(t+1), V(t) | ... -> item of K(t) = 2t - 1 (t >= 1), of V(t) = 2t + 2, last V = 2n - 1.
          const uint32_t n = (uint32_t)p.n_kv_tiles;
          auto item_k = [&](uint32_t t) { return t == 0 ? 0u : 2u * t - 1u; };
          auto item_v = [&](uint32_t t) { return t + 1 < n ? 2u * t + 2u : 2u * n - 1u; };
          auto item_stage = [&](uint32_t it) { return it % (uint32_t)S::kStages; };
          auto item_par = [&](uint32_t it) { return (it / (uint32_t)S::kStages) & 1u; };
          auto stage_smem = [&](uint32_t st) { return base + S::kKvOff + st * S::kKvBytes; };
          uint32_t qk_t[2] = {0u, 0u}, pv_t[2] = {0u, 0u};
          // Q tile B starts p.stagger cycles after Q tile A: two softmax warps that share a scheduler should not be in their
          // exponential phases at the same time (MUFU is the one unit both saturate), and nothing else pulls them apart
          const long long t_b0 = clock64() + p.stagger;
          while (pv_t[0] < n || pv_t[1] < n) {
            // one round: the four conditions are probed back to back (independent ~40-cycle barrier reads), then every
            // group found ready is issued. A condition that was true stays true, so a stale positive is still valid.
            bool rq[2], rp[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              const uint32_t tq = qk_t[x], tp = pv_t[x];
              const bool sf = mbar_test_wait(s_free(x), (tq - 1u) & 1u);
              const bool pr = mbar_test_wait(p_ready(x), tp & 1u);
              rq[x] = tq < n && (tq == 0 ? (x == 0 || clock64() >= t_b0) : sf);
              rp[x] = tp < tq && pr;
            }
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              // ---- QK_X(t): S_X(t-1) is in the softmax warps' registers, K(t) landed ----
              if (rq[x]) {
                const uint32_t tq = qk_t[x];
                const uint32_t it = item_k(tq);
                if (mbar_test_wait(kv_full(item_stage(it)), item_par(it))) {
                  tc_fence_after();
                  trace_ev(p, tr, 2 + x, 0, tq);
                  issue_qk(x, stage_smem(item_stage(it)), no_probe);
                  commit(s_full(x));
                  if (qk_t[x ^ 1] > tq) commit(kv_empty(item_stage(it)));  // the other Q tile used K(t) already: slot free
                  trace_ev(p, tr, 2 + x, 1, tq);
                  qk_t[x] = tq + 1;
                }
              }
              // ---- PV_X(t): P_X(t) written, V(t) landed ----
              if (rp[x]) {
                const uint32_t tp = pv_t[x];
                const uint32_t it = item_v(tp);
                if (mbar_test_wait(kv_full(item_stage(it)), item_par(it))) {
                  tc_fence_after();
                  trace_ev(p, tr, 2 + x, 2, tp);
                  issue_pv(x, stage_smem(item_stage(it)), tp != 0, no_probe);
                  commit(o_done(x));
                  if (pv_t[x ^ 1] > tp) commit(kv_empty(item_stage(it)));
                  trace_ev(p, tr, 2 + x, 3, tp);
                  pv_t[x] = tp + 1;
                }
              }
            }
          }
        } else if (DEC) {
          // P travels through shared memory (or spare TMEM columns), so S_X is free again as soon as the softmax warps hold it in
          // registers: QK_X(t+1) is issued ahead of PV_X(t) and the only per-tile dependency chain left is
          // the softmax itself. Per active tile t the groups go QK_A(t+1), PV_A(t), QK_B(t+1), PV_B(t); ring
          // items (order of first use): K(0) | K(t+1), V(t) | ... While a group is being issued the barriers
          // of the NEXT group are probed without blocking; only if they have not completed by the end of the
          // group does the thread fall back to a blocking wait (a real dependency stall).
          auto stage_smem = [&](uint32_t st) { return base + S::kKvOff + st * S::kKvBytes; };
          auto ring_next = [&](uint32_t& st, uint32_t& par) {
            if (++st == (uint32_t)S::kStages) {
              st = 0;
              par ^= 1u;
            }
          };
          issue_qk(0, stage_smem(0), no_probe);
          commit(s_full(0));
          issue_qk(1, stage_smem(0), no_probe);
          commit(s_full(1));
          commit(kv_empty(0));
          uint32_t st = 1u % S::kStages, par = 0u;  // ring cursor: item 1 = K(1) (or V(0) if there is no tile 1)
          uint32_t t = 0;
          bool ok = false;  // "the barriers of the group about to be issued were seen complete by a probe"
          int jn = next_active(j + 1);
          while (true) {
            const bool has_next = jn < p.n_kv_tiles;
            const int jn2 = has_next ? next_active(jn + 1) : jn;
            const bool has_next2 = has_next && jn2 < p.n_kv_tiles;
            const uint32_t ph = t & 1u;
            const uint32_t sK = st, pK = par;  // K(t+1), if any
            uint32_t sV = st, pV = par;        // V(t)
            if (has_next) ring_next(sV, pV);
            uint32_t sK2 = sV, pK2 = pV;       // K(t+2), if any
            ring_next(sK2, pK2);
            // ---- QK_A(t+1): K(t+1) landed, S_A(t) in registers ----
            if (has_next) {
              trace_ev(p, tr, 2, 4, t);
              if (!ok) {
                mbar_wait(kv_full(sK), pK);
                trace_ev(p, tr, 2, 5, t);
                mbar_wait(s_free(0), ph);
              }
              tc_fence_after();
              trace_ev(p, tr, 2, 6, t);
              ok = false;
              trace_ev(p, tr, 2, 0, t);
              issue_qk(0, stage_smem(sK), [&] { if (!ok) ok = mbar_test_wait(kv_full(sV), pV) && mbar_test_wait(p_ready(0), ph); });
              commit(s_full(0));
              trace_ev(p, tr, 2, 1, t);
            } else {
              ok = false;
            }
            // ---- PV_A(t): V(t) landed, P_A(t) written ----
            if (!ok) {
              mbar_wait(kv_full(sV), pV);
              mbar_wait(p_ready(0), ph);
            }
            tc_fence_after();
            ok = false;
            trace_ev(p, tr, 2, 2, t);
            if (has_next) {
              issue_pv(0, stage_smem(sV), t != 0, [&] { if (!ok) ok = mbar_test_wait(s_free(1), ph); });
            } else {
              issue_pv(0, stage_smem(sV), t != 0, [&] { if (!ok) ok = mbar_test_wait(p_ready(1), ph); });
            }
            commit(o_done(0));
            trace_ev(p, tr, 2, 3, t);
            // ---- QK_B(t+1) ----
            if (has_next) {
              if (!ok) mbar_wait(s_free(1), ph);
              tc_fence_after();
              ok = false;
              trace_ev(p, tr, 3, 0, t);
              issue_qk(1, stage_smem(sK), [&] { if (!ok) ok = mbar_test_wait(p_ready(1), ph); });
              commit(s_full(1));
              commit(kv_empty(sK));
              trace_ev(p, tr, 3, 1, t);
            }
            // ---- PV_B(t) ----
            if (!ok) mbar_wait(p_ready(1), ph);
            tc_fence_after();
            ok = false;
            trace_ev(p, tr, 3, 2, t);
            if (has_next2) {
              issue_pv(1, stage_smem(sV), t != 0,
                       [&] { if (!ok) ok = mbar_test_wait(kv_full(sK2), pK2) && mbar_test_wait(s_free(0), ph ^ 1u); });
            } else {
              issue_pv(1, stage_smem(sV), t != 0, no_probe);
            }
            commit(o_done(1));
            commit(kv_empty(sV));
            trace_ev(p, tr, 3, 3, t);
            if (!has_next) break;
            st = sK2;
            par = pK2;
            j = jn;
            jn = jn2;
            ++t;
          }
        } else {
        issue_qk(0, stage_addr(0), no_probe);
        commit(s_full(0));
        issue_qk(1, stage_addr(0), no_probe);
        commit(s_full(1));
        commit(kv_empty(0));
        uint32_t t = 0;
        // active tiles are numbered t = 0,1,...; K(t) is ring slot 2t, V(t) is ring slot 2t+1
        while (true) {
          const int jn = next_active(j + 1);
          const bool has_next = jn < p.n_kv_tiles;
          const uint32_t uV = 2 * t + 1, uKn = 2 * t + 2;
          const uint32_t ph = t & 1u;
          // both operand tiles of this half-iteration were requested more than an iteration ago: take their
          // (already satisfied, but ~100-cycle) barrier waits BEFORE blocking on the softmax, so that PV_A
          // and QK_A go out back to back once P_A arrives
          wait_full(uV);
          if (has_next) wait_full(uKn);
          trace_ev(p, tr, 2, 0, t);
          mbar_wait(p_ready(0), ph);
          trace_ev(p, tr, 2, 1, t);
          tc_fence_after();
          issue_pv(0, stage_addr(uV), t != 0, no_probe);
          trace_ev(p, tr, 2, 4, t);
          commit(o_done(0));
          if (has_next) {
            trace_ev(p, tr, 2, 5, t);
            issue_qk(0, stage_addr(uKn), no_probe);
            commit(s_full(0));
          }
          trace_ev(p, tr, 2, 2, t);
          mbar_wait(p_ready(1), ph);
          trace_ev(p, tr, 2, 3, t);
          tc_fence_after();
          issue_pv(1, stage_addr(uV), t != 0, no_probe);
          commit(o_done(1));
          commit(kv_empty(uV % S::kStages));
          if (!has_next) break;
          issue_qk(1, stage_addr(uKn), no_probe);
          commit(s_full(1));
          commit(kv_empty(uKn % S::kStages));
          j = jn;
          ++t;
        }
        }
      }
    }
    __syncwarp();
  }
  } else {
    // ===================== softmax / correction / epilogue =====================
    reg_alloc<208>();
    const int x = warp >> 2;              // Q tile of this warpgroup
    const int lane_group = warp & 3;      // TMEM lanes [32*lane_group, +32)
    const int row_in_tile = lane_group * 32 + lane;
    const int row = q0 + x * kQTile + row_in_tile;  // global query index
    const uint32_t lane_off = (uint32_t)(lane_group * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + (uint32_t)(x * 128);
    const uint32_t tO = tmem_base + lane_off + 256u + (uint32_t)(x * 128);
    const uint32_t tP = PT ? tO + 64u : tS;  // where P goes in TMEM (unless it goes to shared memory)
    // this thread's row of the K-major, 128B-swizzled P tile in shared memory (PS)
    const uint32_t p_row = base + S::kPOff + (uint32_t)x * S::kPBytes + (uint32_t)row_in_tile * 128u;
    const uint32_t p_sw = (uint32_t)(row_in_tile & 7);
    const int8_t* mask_row = nullptr;
    if (has_mask) mask_row = mask_bh + (int64_t)min(row / p.mask_bq, p.nbq - 1) * p.nbk;

    float m_run = -INFINITY;  // running max, scaled-log2 domain
    float l_run = 0.f;
    uint32_t t = 0;
    for (int j = next_active(0); j < p.n_kv_tiles; j = next_active(j + 1)) {
      // (TRACE: roles 0/1 = this CTA's warp 0 of Q tile A/B for CTA 0; roles 4/5 = the same warps of CTA 1, the peer of a pair)
      const bool tr = TRACE && p.trace != nullptr && blockIdx.x <= 1 && blockIdx.y == 0 && blockIdx.z == 0 && (warp & 3) == 0 && lane == 0;
      const int xr = x + 4 * (int)blockIdx.x;
      trace_ev(p, tr, xr, 0, t);
      mbar_wait(s_full(x), t & 1u);
      trace_ev(p, tr, xr, 1, t);
      tc_fence_after();
      // PS: has PV_X(t-1) finished reading the P tile? Probed here, needed only at the first store of P
      bool p_free = !DEC || t == 0;
      if (DEC && t > 0) p_free = mbar_test_wait(o_done(x), (t - 1) & 1u);
      const int valid = p.Sk - j * kKvTile;  // keys of this tile inside the sequence
      bool seg0 = true, seg1 = true;
      if (has_mask) {
        const int kb0 = min((j * kKvTile) / p.mask_bk, p.nbk - 1);
        const int kb1 = min((j * kKvTile + 64) / p.mask_bk, p.nbk - 1);
        seg0 = mask_row[kb0] != 0;
        seg1 = mask_row[kb1] != 0;
      }
      // only tiles that are ragged at the sequence end or partially masked for this row need the -inf pass
      const bool masked_tile = valid < kKvTile || !seg0 || !seg1;
      // this warp's 32 rows do not attend to any key of the tile (it is active for other query blocks of the
      // CTA / CTA pair): P = 0, the running max and sum stay as they are -- no S read, no exponentials
      const bool dead = has_mask && __all_sync(0xffffffffu, !seg0 && !seg1);
      bool tmem_dirty = false;
      if (dead) {
        if (DEC) {
          tc_fence_before();
          if (CG == 2) {
            __syncwarp();
            if (lane == 0) arrive_lead(s_free(x));
          } else {
            mbar_arrive(s_free(x));
          }
          if (!p_free) mbar_wait(o_done(x), (t - 1) & 1u);
        }
        if (PS) {
#pragma unroll
          for (int q = 0; q < kKvTile * ES / 16; ++q)
            sts128(p_row + (uint32_t)((q >> 3) * (kQTile * 128)) + ((((uint32_t)(q & 7)) ^ p_sw) << 4), 0u, 0u, 0u, 0u);
        } else {
          uint32_t z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
          for (int c = 0; c < kKvTile * ES / 64; ++c) tmem_st_32x16(tP + (uint32_t)(c * 16), z);
        }
      } else {
      // -inf on keys beyond the sequence end and on masked 64-key segments
      auto apply_mask = [&](uint32_t(&r)[32], int c) {
        if (masked_tile) {
          const bool seg = c < 2 ? seg0 : seg1;
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = (seg && (c * 32 + i) < valid) ? r[i] : 0xff800000u;
        }
      };

      // ---- S row -> registers (one TMEM read; the softmax warpgroups run at 208 registers) ----
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld_32x32(tS, s0);
      tmem_ld_32x32(tS + 32u, s1);
      tmem_ld_32x32(tS + 64u, s2);
      tmem_ld_32x32(tS + 96u, s3);
      tmem_ld_wait();
      if (DEC) {
        // S_X is in registers: the MMA warp may overwrite it with the next tile's scores right away
        tc_fence_before();
        if (CG == 2) {
          __syncwarp();
          if (lane == 0) arrive_lead(s_free(x));
        } else {
          mbar_arrive(s_free(x));
        }
      }
      trace_ev(p, tr, xr, 2, t);
      apply_mask(s0, 0);
      apply_mask(s1, 1);
      apply_mask(s2, 2);
      apply_mask(s3, 3);
#ifdef FDM_ATTN_CHEAT_NOMAX
      const float mx = t == 0 ? fmaxf(fmax3(max32(s0), max32(s1), max32(s2)), max32(s3)) + 6.f : -INFINITY;   // (upper-bound experiment only)
#else
      const float mx = fmaxf(fmax3(max32(s0), max32(s1), max32(s2)), max32(s3));
#endif
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const bool need = m_new > m_run + kRescaleThreshold;  // first finite max always triggers
      if (__any_sync(0xffffffffu, need)) {
        const float m_next = need ? m_new : m_run;
        const float alpha = (m_run == -INFINITY) ? 0.f : ex2(m_run - m_next);
        l_run *= alpha;
        m_run = m_next;
        if (t > 0) {
          // O_X may only be touched between PV_X(t-1) and PV_X(t)
          tmem_dirty = true;
          mbar_wait(o_done(x), (t - 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tO + (uint32_t)(c * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32(tO + (uint32_t)(c * 32), r);
          }
        }
      }
      trace_ev(p, tr, xr, 3, t);
      // the P tile (shared memory / spare TMEM columns) is read by PV_X(t-1) until o_done(x) completes its phase.
      // LS: all exponentials of the tile are computed and packed into registers first (64 registers replace the 128 of
      // S as they are consumed) and the wait comes just before the burst of stores -- PV_X(t-1) has ~the whole softmax
      // of tile t to finish instead of its first third, so the single P buffer stops serialising PV(t-1) and exp(t).
      if (!LS && DEC && !p_free) mbar_wait(o_done(x), (t - 1) & 1u);
      // ---- P = exp2(S*scale - m): over the first columns of S_X in TMEM, or into the P tile in smem ----
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
      const uint64_t scale2 = f2(p.scale_log2, p.scale_log2), negm2 = f2(neg_m, neg_m);
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};  // 4 independent packed partial row sums
      constexpr int kPW = DT == kDtE4M3 ? 8 : 16;  // packed 32-bit words per 32 keys
      auto compute = [&](uint32_t(&r)[32], uint32_t(&pk)[kPW]) {
        float pv[DT == kDtE4M3 ? 32 : 1];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t y = ffma2(f2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), scale2, negm2);
          float p0, p1;
          if (2 * i < EMU) {
            ex2_emulated_pair(y, p0, p1);
          } else {
            float y0, y1;
            unf2(y, y0, y1);
            p0 = ex2(y0);
            p1 = ex2(y1);
          }
          acc[i & 3] = fadd2(acc[i & 3], f2(p0, p1));
          if constexpr (DT == kDtE4M3) {
            pv[2 * i] = p0;
            pv[2 * i + 1] = p1;
          } else {
            pk[i] = pack2<DT>(p0, p1);
          }
        }
        if constexpr (DT == kDtE4M3) {
          // P -> e4m3, unscaled (the reference's fp8 semantics), 4 keys per 32-bit word
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = cvt_e4m3x4(pv[4 * i], pv[4 * i + 1], pv[4 * i + 2], pv[4 * i + 3]);
        }
      };
      auto store = [&](const uint32_t(&pk)[kPW], int c) {
        if constexpr (DT == kDtE4M3) {
          if (PS) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
              sts128(p_row + ((((uint32_t)(c * 2 + q)) ^ p_sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          } else {
            tmem_st_32x8(tP + (uint32_t)(c * 8), reinterpret_cast<const uint32_t(&)[8]>(pk));
          }
        } else if (PS) {
          // 32 keys = 64 bytes = four 16-byte chunks of this row in panel c/2 (64 keys per 128-byte row)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(p_row + (uint32_t)((c >> 1) * (kQTile * 128)) + ((((uint32_t)((c & 1) * 4 + q)) ^ p_sw) << 4),
                   pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        } else {
          tmem_st_32x16(tP + (uint32_t)(c * 16), reinterpret_cast<const uint32_t(&)[16]>(pk));
        }
      };
      if constexpr (LS) {
        uint32_t pk0[kPW], pk1[kPW], pk2[kPW], pk3[kPW];
        compute(s0, pk0);
        compute(s1, pk1);
        compute(s2, pk2);
        compute(s3, pk3);
        trace_ev(p, tr, xr, 7, t);
        if (DEC && !p_free) mbar_wait(o_done(x), (t - 1) & 1u);
        store(pk0, 0);
        store(pk1, 1);
        store(pk2, 2);
        store(pk3, 3);
      } else {
        uint32_t pk[kPW];
        compute(s0, pk);
        store(pk, 0);
        compute(s1, pk);
        store(pk, 1);
        compute(s2, pk);
        store(pk, 2);
        compute(s3, pk);
        store(pk, 3);
      }
      {
        float a0, a1, b0, b1;
        unf2(fadd2(acc[0], acc[1]), a0, a1);
        unf2(fadd2(acc[2], acc[3]), b0, b1);
        // a row that has seen no unmasked key yet must keep l = 0: the emulated exp2 returns 2^-126, not 0,
        // for -inf inputs (harmless next to real probabilities, but not as the only terms of the sum)
        l_run += (m_run == -INFINITY) ? 0.f : (a0 + a1) + (b0 + b1);
      }
      }  // !dead
      trace_ev(p, tr, xr, 4, t);
      if (PS) fence_proxy_async_smem();  // generic-proxy stores of P -> visible to the tensor core's reads
      if (!PS || tmem_dirty) tmem_st_wait();  // P in TMEM, or O rescaled through TMEM
      tc_fence_before();
      if (CG == 2) {
        __syncwarp();
        if (lane == 0) arrive_lead(p_ready(x));
      } else {
        mbar_arrive(p_ready(x));
      }
      trace_ev(p, tr, xr, 5, t);
      // (role 6: arrival time of each of the leader CTA's eight softmax warps, event = warp)
      trace_ev(p, TRACE && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0, 6, warp, t);
      if (TRACE && p.trace != nullptr && blockIdx.x <= 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && t < kTraceTiles)
        atomicMax((unsigned long long*)&p.trace[(xr * 8 + 6) * kTraceTiles + t], (unsigned long long)clock64());
      ++t;
    }

    // ---- epilogue: O / l -> global ----
    if (life) p.trace[(3 * 8 + 7) * kTraceTiles + 2] = clock64();
    const bool row_ok = row < p.Sq;
    uint16_t* out_row;
    if constexpr (SCATTER) {
      // (batch 1) the row's owner and its row there; rows beyond Sq are never stored
      const int dest = min(row / sp.rows_per_peer, 7);
      out_row = reinterpret_cast<uint16_t*>(sp.o_peer[dest]) + (int64_t)(row - dest * sp.rows_per_peer) * p.o_ts + (int64_t)h * HD;
    } else {
      out_row = reinterpret_cast<uint16_t*>(p.o) + (int64_t)b * p.o_bs + (int64_t)row * p.o_ts + (int64_t)h * HD;
    }
    if (t > 0) {
      mbar_wait(o_done(x), (t - 1) & 1u);
      tc_fence_after();
    }
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      if (t > 0) {
        tmem_ld_32x32(tO + (uint32_t)(c * 32), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (row_ok) {
        U128 o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          o[q].x = pack2<DT>(__uint_as_float(r[q * 8 + 0]) * inv, __uint_as_float(r[q * 8 + 1]) * inv);
          o[q].y = pack2<DT>(__uint_as_float(r[q * 8 + 2]) * inv, __uint_as_float(r[q * 8 + 3]) * inv);
          o[q].z = pack2<DT>(__uint_as_float(r[q * 8 + 4]) * inv, __uint_as_float(r[q * 8 + 5]) * inv);
          o[q].w = pack2<DT>(__uint_as_float(r[q * 8 + 6]) * inv, __uint_as_float(r[q * 8 + 7]) * inv);
        }
        if (p.o_vec32) {  // 32-byte aligned rows: 256-bit stores (half the requests, whole sectors)
          stg256(out_row + c * 32, o[0], o[1]);
          stg256(out_row + c * 32 + 16, o[2], o[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) stg128(out_row + c * 32 + q * 8, o[q]);
        }
      }
    }
  }

  if (life) p.trace[(3 * 8 + 7) * kTraceTiles + 3] = clock64();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();  // neither CTA's shared memory / TMEM goes away while the pair is still working
  if (life) p.trace[(3 * 8 + 7) * kTraceTiles + 4] = clock64();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

template <int HD, int DT, int EMUX, bool PS, int CG, bool MASKED, bool TRACE>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  attn_fwd_body<HD, DT, EMUX, PS, CG, MASKED, TRACE, false>(tmap_q, tmap_k, tmap_v, p, ScatterParams{});
}
template <int HD, int DT, int EMUX, bool PS, int CG, bool MASKED>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_scatter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const AttnParams p, const ScatterParams sp) {
  attn_fwd_body<HD, DT, EMUX, PS, CG, MASKED, false, true>(tmap_q, tmap_k, tmap_v, p, sp);
}

template <int HD, int DT, int EMU, bool PS, int CG, bool MASKED, bool TRACE, bool SCATTER = false>
static int launch_attn_k(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         const AttnParams& p, cudaStream_t st, const ScatterParams* sp = nullptr) {
  using S = AttnSmem<HD, DT == kDtE4M3 ? 1 : 2, PS, CG>;
  static std::atomic<bool> attr_set[64];  // zero-initialised; setting the attribute twice is harmless
  int dev = 0;
  FDM_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    if constexpr (SCATTER)
      FDM_CUDA(cudaFuncSetAttribute(attn_fwd_scatter_kernel<HD, DT, EMU, PS, CG, MASKED>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    else
      FDM_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD, DT, EMU, PS, CG, MASKED, TRACE>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set[dev].store(true, std::memory_order_release);
  }
  const unsigned nq = (unsigned)((p.Sq + 2 * kQTile - 1) / (2 * kQTile));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((nq + CG - 1) / CG * CG, (unsigned)p.H, (unsigned)p.B);  // whole CTA pairs
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if constexpr (SCATTER)
    FDM_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_scatter_kernel<HD, DT, EMU, PS, CG, MASKED>, tq, tk, tv, p, *sp));
  else
    FDM_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_kernel<HD, DT, EMU, PS, CG, MASKED, TRACE>, tq, tk, tv, p));
  FDM_LAUNCH_CHECK("attn_fwd kernel launch");
  return FDM_OK;
}

template <int HD, int DT, int EMU, bool PS, int CG>
static int launch_attn_p(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         const AttnParams& p, cudaStream_t st, const ScatterParams* sp) {
  if (p.trace != nullptr) {
    // the timeline build exists for the default dense bf16 hd-128 configurations only
    if constexpr (HD == 128 && DT == kDtBF16 && (EMU & 63) == 4) {
      if (p.mask == nullptr) return launch_attn_k<HD, DT, EMU, PS, CG, false, true>(tq, tk, tv, p, st);
    }
  }
  if (sp != nullptr) {
    // epilogue scatters the rows to their owners' buffers (Ulysses): built for the Wan / Qwen case, hd 128, bf16
    if constexpr (HD == 128 && DT == kDtBF16 && (EMU & 63) == 4) {
      if constexpr (CG == 1) {
        if (p.mask != nullptr) return launch_attn_k<HD, DT, EMU, PS, CG, true, false, true>(tq, tk, tv, p, st, sp);
      }
      if (p.mask == nullptr) return launch_attn_k<HD, DT, EMU, PS, CG, false, false, true>(tq, tk, tv, p, st, sp);
    }
    set_error("attn: the scattering epilogue is built for head_dim 128, bf16, default exp2 split");
    return FDM_ERR_UNSUPPORTED;
  }
  if constexpr (CG == 1) {  // block-sparse calls never take the CTA-pair kernel (attn_use_pair)
    if (p.mask != nullptr) return launch_attn_k<HD, DT, EMU, PS, CG, true, false>(tq, tk, tv, p, st);
  } else if (p.mask != nullptr) {
    set_error("attn: internal error, block mask routed to the CTA-pair kernel");
    return FDM_ERR_UNSUPPORTED;
  }
  return launch_attn_k<HD, DT, EMU, PS, CG, false, false>(tq, tk, tv, p, st);
}

// experiment knob: FDM_ATTN_CG=1 runs single CTAs (P in TMEM over S) instead of CTA pairs (P through shared
// memory) for hd 128 with 16-bit operands
static int attn_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int attn_env_stagger() {
  static const int v = attn_env("FDM_ATTN_STAGGER", 1200);
  return v;
}
static bool attn_pair_setting() {
  static int v = attn_env("FDM_ATTN_CG", 2);
  return v == 2;
}
// CTA pairs pay off once the K/V loop is long enough to amortise the cluster set-up and there are at least
// two 256-row query blocks to pair (cross-attention onto 512 text tokens stays on single CTAs)
static bool attn_use_pair(int64_t Sq, int64_t Sk, bool masked) {
  // block-sparse calls stay on single CTAs: a pair shares one active-tile list over 512 query rows, which for
  // band-shaped (radial) masks keeps 20-25 % more tiles than 256-row CTAs do (measured 1.85x vs 2.70x speed-up
  // over dense at 24 % block density)
  return attn_pair_setting() && !masked && Sk >= 8 * kKvTile && Sq > 2 * kQTile;
}
template <int HD, int DT, int EMU>
static int launch_attn_e(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         const AttnParams& p, cudaStream_t st, const ScatterParams* sp) {
  if constexpr (HD == 128 && DT != kDtE4M3) {
    if (attn_use_pair(p.Sq, p.Sk, p.mask != nullptr)) return launch_attn_p<HD, DT, EMU, true, 2>(tq, tk, tv, p, st, sp);
  }
  if constexpr (DT == kDtE4M3) {
    // fp8 operands: Q, K, V and P tiles are 16 KB, so a single CTA has room for P in shared memory next to a 6-stage ring
    // (bf16 does not: 3 stages, which starve). S_X is then free as soon as the softmax warps hold it and QK_X(t+1) goes out
    // ahead of PV_X(t), as in the CTA-pair kernel. Measured: 8704^2 x 24 heads 1288 -> 1621 TFLOP/s, 80640^2 x 4 heads
    // 1847 -> 1784: dense calls of up to 16384 keys take it (FDM_ATTN_FP8_PS=0: never, =2: always)
    static const int ps = attn_env("FDM_ATTN_FP8_PS", 1);
    if (ps && p.mask == nullptr && sp == nullptr && p.Sk >= 4 * kKvTile && (ps == 2 || p.Sk <= 16384)) return launch_attn_p<HD, DT, EMU, true, 1>(tq, tk, tv, p, st, sp);
  }
  return launch_attn_p<HD, DT, EMU, false, 1>(tq, tk, tv, p, st, sp);
}

// how many of every 32 exponentials run on the FMA pipe instead of MUFU (tuning knob; the default is
// the measured optimum, FDM_ATTN_EMU overrides it for experiments)
static int attn_emu_setting(int hd, bool fp8) {
  static const int v = attn_env("FDM_ATTN_EMU", -1);
  if (v >= 0) return v;
  // fp8 operands: the MMAs take half the time, so the softmax is further from hiding behind them and more of its
  // exponentials pay off on the FMA pipe (8704^2 x 24: 1134 -> 1561 TFLOP/s, 80640^2 x 4: 1565 -> 1806 with 12 instead of 4)
  if (fp8) return 12;
  return hd == 128 ? 4 : 12;
}

// late store of P (see the softmax loop); FDM_ATTN_LS=0 restores the store-as-you-go order for experiments
static bool attn_late_store() {
  static const int v = attn_env("FDM_ATTN_LS", 1);
  return v != 0;
}

template <int HD, int DT>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const AttnParams& p, cudaStream_t st, const ScatterParams* sp = nullptr) {
  const int emu = attn_emu_setting(HD, DT == kDtE4M3);
  if (attn_late_store()) {
    if (emu <= 0) return launch_attn_e<HD, DT, 64 + 0>(tq, tk, tv, p, st, sp);
    if (emu <= 4) return launch_attn_e<HD, DT, 64 + 4>(tq, tk, tv, p, st, sp);
    if (emu <= 8) return launch_attn_e<HD, DT, 64 + 8>(tq, tk, tv, p, st, sp);
    if (emu <= 12) return launch_attn_e<HD, DT, 64 + 12>(tq, tk, tv, p, st, sp);
    return launch_attn_e<HD, DT, 64 + 16>(tq, tk, tv, p, st, sp);
  }
  if (emu <= 0) return launch_attn_e<HD, DT, 0>(tq, tk, tv, p, st, sp);
  if (emu <= 4) return launch_attn_e<HD, DT, 4>(tq, tk, tv, p, st, sp);
  return launch_attn_e<HD, DT, 8>(tq, tk, tv, p, st, sp);
}

static int make_qkv_tmap(CUtensorMap* out, const void* ptr, int64_t B, int64_t S, int H, int hd,
                         int64_t batch_stride, int64_t token_stride, int es, int box_rows = 128) {
  // dims innermost first: d, head, token, batch; one box = box_rows rows x 128 bytes
  uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)hd * es, (uint64_t)token_stride * es, (uint64_t)batch_stride * es};
  uint32_t box[4] = {(uint32_t)(128 / es), 1, (uint32_t)box_rows, 1};
  return make_tmap(out, es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, ptr,
                   dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace fdm

using namespace fdm;

static long long* g_attn_trace = nullptr;
extern "C" int fdm_debug_set_attn_trace(void* device_buffer) {
  g_attn_trace = (long long*)device_buffer;  // 7 roles x 8 events x 64 tiles x int64, or NULL to disable
  return FDM_OK;
}

namespace fdm {
// csrc/attention_persistent.cu: the same arguments, dense calls only
int attn_fwd_persistent(const void* q, const void* k, const void* v, void* o, const int8_t* block_mask, int64_t B,
                        int64_t Sq, int64_t Sk, int H, int hd, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts,
                        int64_t v_bs, int64_t v_ts, int64_t o_bs, int64_t o_ts, int mask_bq, int mask_bk, float scale,
                        int qkv_dtype, void* stream);
// Dense calls with at most this many KV tiles go to the persistent kernel (FDM_ATTN_PERSIST_TILES; 0 = never): its
// item loop hides the per-CTA set-up / first loads / epilogue (~10-20 us), which is most of a 4-tile CTA's life; on
// longer calls this file's one-item kernel is faster (see the header of attention_persistent.cu).
static int attn_persist_tiles() {
  static const int v = attn_env("FDM_ATTN_PERSIST_TILES", 16);
  return v;
}
}  // namespace fdm

static int attn_fwd_impl(const void* q, const void* k, const void* v, void* o, void* const* o_peers, int n_peers,
                         int64_t rows_per_peer, const int8_t* block_mask, int64_t B, int64_t Sq, int64_t Sk, int H, int hd,
                         int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs,
                         int64_t v_ts, int64_t o_bs, int64_t o_ts, int mask_bq, int mask_bk,
                         float scale, int qkv_dtype, void* stream) {
  if (o_peers == nullptr && block_mask == nullptr && g_attn_trace == nullptr && Sk > 0 &&
      (Sk + kKvTile - 1) / kKvTile <= attn_persist_tiles())
    return attn_fwd_persistent(q, k, v, o, block_mask, B, Sq, Sk, H, hd, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts,
                               mask_bq, mask_bk, scale, qkv_dtype, stream);
  int rc = require_sm100();
  if (rc) return rc;
  FDM_REQUIRE(B >= 0 && Sq >= 0 && Sk >= 0 && H > 0, "attn: bad shape");
  if (B == 0 || Sq == 0) return FDM_OK;
  FDM_REQUIRE(q && k && v && (o || o_peers), "attn: null pointer");
  FDM_REQUIRE(hd == 64 || hd == 128, "attn: head_dim %d unsupported (64 or 128)", hd);
  FDM_REQUIRE(qkv_dtype == FDM_BF16 || qkv_dtype == FDM_F16 || qkv_dtype == FDM_E4M3,
              "attn: q/k/v dtype must be bf16, f16 or e4m3");
  const int es = qkv_dtype == FDM_E4M3 ? 1 : 2;
  if (es == 1 && hd != 128) {
    set_error("attn: fp8 q/k/v is built for head_dim 128 only");
    return FDM_ERR_UNSUPPORTED;
  }
  FDM_REQUIRE(scale > 0.f, "attn: scale must be positive");
  FDM_REQUIRE(Sk > 0, "attn: empty key sequence");
  FDM_REQUIRE(Sq < (1LL << 31) && Sk <= (int64_t)kMaxKvTiles * kKvTile && H < 65536 && B < 65536,
              "attn: sequence too long (Sk <= %d)", kMaxKvTiles * kKvTile);
  for (int64_t s : {q_ts, k_ts, v_ts, q_bs, k_bs, v_bs})
    FDM_REQUIRE((s * es) % 16 == 0, "attn: q/k/v strides must be multiples of 16 bytes");
  FDM_REQUIRE(o_ts % 8 == 0 && o_bs % 8 == 0, "attn: output strides must be multiples of 8 elements");
  FDM_REQUIRE((uintptr_t)q % 16 == 0 && (uintptr_t)k % 16 == 0 && (uintptr_t)v % 16 == 0 &&
                  (uintptr_t)o % 16 == 0,
              "attn: pointers must be 16-byte aligned");
  FDM_REQUIRE(q_ts >= (int64_t)H * hd && k_ts >= (int64_t)H * hd && v_ts >= (int64_t)H * hd,
              "attn: token stride smaller than H*hd");
  FDM_REQUIRE(o_ts >= (int64_t)H * hd, "attn: output token stride smaller than H*hd");
  AttnParams p;
  p.o = o;
  p.o_bs = o_bs;
  p.o_ts = o_ts;
  p.o_vec32 = (o_ts % 16 == 0 && o_bs % 16 == 0 && (uintptr_t)o % 32 == 0 && (hd * 2) % 32 == 0) ? 1 : 0;
  p.mask = block_mask;
  p.B = (int)B;
  p.H = H;
  p.Sq = (int)Sq;
  p.Sk = (int)Sk;
  p.n_kv_tiles = (int)((Sk + kKvTile - 1) / kKvTile);
  p.mask_bq = mask_bq;
  p.mask_bk = mask_bk;
  p.nbq = p.nbk = 0;
  if (block_mask) {
    FDM_REQUIRE((mask_bq == 64 || mask_bq == 128) && (mask_bk == 64 || mask_bk == 128),
                "attn: mask block sizes must be 64 or 128");
    p.nbq = (int)((Sq + mask_bq - 1) / mask_bq);
    p.nbk = (int)((Sk + mask_bk - 1) / mask_bk);
  }
  p.scale_log2 = scale * 1.4426950408889634f;
  p.stagger = attn_env_stagger();
  p.trace = g_attn_trace;
  ScatterParams sc;
  sc.rows_per_peer = 0;
  for (int i = 0; i < 8; ++i) sc.o_peer[i] = nullptr;
  if (o_peers != nullptr) {
    FDM_REQUIRE(B == 1 && n_peers >= 1 && n_peers <= 8 && rows_per_peer > 0 && rows_per_peer * n_peers >= Sq &&
                    rows_per_peer < (1LL << 31),
                "attn: scatter needs batch 1, 1..8 peers and rows_per_peer * peers >= Sq");
    for (int i = 0; i < n_peers; ++i) {
      FDM_REQUIRE(o_peers[i] != nullptr && (uintptr_t)o_peers[i] % 16 == 0, "attn: peer output pointers must be 16-byte aligned");
      sc.o_peer[i] = o_peers[i];
      p.o_vec32 = p.o_vec32 && ((uintptr_t)o_peers[i] % 32 == 0);
    }
    for (int i = n_peers; i < 8; ++i) sc.o_peer[i] = o_peers[n_peers - 1];
    sc.rows_per_peer = (int)rows_per_peer;
  }
  const ScatterParams* sp = o_peers != nullptr ? &sc : nullptr;
  CUtensorMap tq, tk, tv;
  // batch stride of a single-batch tensor is irrelevant but must still be a legal stride
  if (B == 1) {
    q_bs = Sq * q_ts;
    k_bs = Sk * k_ts;
    v_bs = Sk * v_ts;
  }
  rc = make_qkv_tmap(&tq, q, B, Sq, H, hd, q_bs, q_ts, es);
  if (rc) return rc;
  // a CTA pair splits every K tile by keys: 64-row boxes
  const bool pair = hd == 128 && es == 2 && attn_use_pair(Sq, Sk, block_mask != nullptr);
  rc = make_qkv_tmap(&tk, k, B, Sk, H, hd, k_bs, k_ts, es, pair ? 64 : 128);
  if (rc) return rc;
  rc = make_qkv_tmap(&tv, v, B, Sk, H, hd, v_bs, v_ts, es);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (qkv_dtype == FDM_E4M3) return launch_attn<128, kDtE4M3>(tq, tk, tv, p, st, sp);
  const bool f16 = qkv_dtype == FDM_F16;
  if (hd == 128) return f16 ? launch_attn<128, kDtF16>(tq, tk, tv, p, st, sp) : launch_attn<128, kDtBF16>(tq, tk, tv, p, st, sp);
  return f16 ? launch_attn<64, kDtF16>(tq, tk, tv, p, st, sp) : launch_attn<64, kDtBF16>(tq, tk, tv, p, st, sp);
}

extern "C" int fdm_attn_fwd(const void* q, const void* k, const void* v, void* o,
                            const int8_t* block_mask, int64_t B, int64_t Sq, int64_t Sk, int H, int hd,
                            int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs,
                            int64_t v_ts, int64_t o_bs, int64_t o_ts, int mask_bq, int mask_bk,
                            float scale, int qkv_dtype, void* stream) {
  return attn_fwd_impl(q, k, v, o, nullptr, 0, 0, block_mask, B, Sq, Sk, H, hd, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts,
                       mask_bq, mask_bk, scale, qkv_dtype, stream);
}

extern "C" int fdm_attn_fwd_scatter(const void* q, const void* k, const void* v, void* const* o_peers, int n_peers,
                                    int64_t rows_per_peer, const int8_t* block_mask, int64_t Sq, int64_t Sk, int H, int hd,
                                    int64_t q_ts, int64_t k_ts, int64_t v_ts, int64_t o_ts, int mask_bq, int mask_bk,
                                    float scale, int qkv_dtype, void* stream) {
  FDM_REQUIRE(o_peers != nullptr && n_peers >= 1, "attn: scatter needs the peers' output pointers");
  return attn_fwd_impl(q, k, v, o_peers[0], o_peers, n_peers, rows_per_peer, block_mask, 1, Sq, Sk, H, hd, Sq * q_ts, q_ts,
                       Sk * k_ts, k_ts, Sk * v_ts, v_ts, rows_per_peer * o_ts, o_ts, mask_bq, mask_bk, scale, qkv_dtype, stream);
}
