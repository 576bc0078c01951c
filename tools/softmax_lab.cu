// Micro-benchmark ("softmax lab"): the per-tile softmax of the attention kernel in isolation -- no MMA, no TMA, no
// barriers -- to find what one SM sub-partition can sustain with ONE and with TWO softmax warps (the kernel runs two: Q
// tiles A and B) and to try loop structures before they go into csrc/attention.cu.
//   S row (128 fp32 columns) <- tcgen05.ld from TMEM, row max, P = exp2(S*scale - m) as bf16, row sum, P -> shared memory
//   (128B-swizzled K-major rows, as the PV MMA reads them).
// Variants (VAR):
//   0  kernel as of round 2: ld all | max | exp (EMU of 32 on the FMA pipe) | store all
//   1  speculative max: exponentials use the running max known BEFORE the tile; the tile's own max is computed in the
//      same instruction stream (checked afterwards; the rare miss would redo the tile)
//   2  1 + chunked TMEM read (wait for the first 32 columns only, the other three loads stay in flight)
//   3  probes: MUFU only / FFMA2 only / MUFU + k FFMA2
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -I include -o build/softmax_lab tools/softmax_lab.cu
#include <cstdio>
#include <cstdlib>
#include "sm100.cuh"
namespace fdm {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -3; }
int require_sm100() { return 0; }
int num_sms() { return 148; }
}
using namespace fdm;
using namespace fdm::sm100;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t f2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unf2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t pk_bf16(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = fmax3(__uint_as_float(r[3 * i]), __uint_as_float(r[3 * i + 1]), __uint_as_float(r[3 * i + 2]));
  m[10] = fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31]));
  const float a = fmax3(m[0], m[1], m[2]), b = fmax3(m[3], m[4], m[5]), c = fmax3(m[6], m[7], m[8]);
  return fmax3(fmax3(a, b, c), m[9], m[10]);
}
__device__ __forceinline__ void ex2_emulated_pair(uint64_t y, float& p0, float& p1) {
  float y0, y1;
  unf2(y, y0, y1);
  y = f2(fmaxf(y0, -126.f), fmaxf(y1, -126.f));
  const uint64_t t = fadd2(y, f2(12582912.f, 12582912.f));
  const uint64_t n = fadd2(t, f2(-12582912.f, -12582912.f));
  const uint64_t f = ffma2(n, f2(-1.f, -1.f), y);
  uint64_t q = ffma2(f, f2(0.05517186224460602f, 0.05517186224460602f), f2(0.2426111400127411f, 0.2426111400127411f));
  q = ffma2(q, f, f2(0.6932609677314758f, 0.6932609677314758f));
  q = ffma2(q, f, f2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  unf2(q, q0, q1);
  unf2(t, t0, t1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// P = exp2(r * scale + negm) for 32 columns: EMU of them on the FMA pipe; packed row sums into acc[4]
template <int EMU>
__device__ __forceinline__ void compute32(const uint32_t (&r)[32], uint32_t (&pk)[16], uint64_t scale2, uint64_t negm2, uint64_t (&acc)[4]) {
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
    pk[i] = pk_bf16(p0, p1);
  }
}
__device__ __forceinline__ void store32(uint32_t p_row, uint32_t p_sw, const uint32_t (&pk)[16], int c) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    sts128(p_row + (uint32_t)((c >> 1) * (128 * 128)) + ((((uint32_t)((c & 1) * 4 + q)) ^ p_sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

template <int VAR, int EMU, bool ST>
__global__ void __launch_bounds__(384, 1) lab(int nwarps, int tiles, float scale_log2, long long* out, float* sink, int mma_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done_warps;
  __shared__ uint64_t mbar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8) tmem_alloc<1>(smem_u32(&tmem_ptr), 512);
  if (threadIdx.x == 0) { done_warps = 0; mbar_init(smem_u32(&mbar), 1); fence_mbar_init(); }
  // operands of the dummy MMAs: zeros (S stays finite); P region [0, 64K) is written by the softmax warps
  for (uint32_t i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) sts128(base + i * 16, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp >= 8) {
    reg_dealloc<88>();
    if (warp == 9 && mma_mode != 0) {
      // the kernel's MMA stream without any dependency: QK_A, PV_A, QK_B, PV_B (8 M128 N128 K16 MMAs each), free-running.
      // mode 1: PV takes P from shared memory (SS, as the CTA-pair kernel); mode 2: PV takes P from TMEM (TS, over S)
      const uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 1);
      const uint32_t q_smem = base + 65536, k_smem = base + 131072, v_smem = base + 163840;
      long long n_mma = 0;
      const long long t0 = clock64();
      const int style = mma_mode >> 2;
      mma_mode &= 3;
      if (mma_mode == 3) {
        if (elect_one()) {
          while (done_warps < nwarps) {
            for (int i = 0; i < 64; ++i)
              if (mbar_try_wait(smem_u32(&mbar), 0)) break;
          }
          out[148 * 8 * 5 + blockIdx.x * 2] = 1; out[148 * 8 * 5 + blockIdx.x * 2 + 1] = 1;
        }
        __syncwarp();
      } else
      if (style == 1) {
        // one thread runs the whole issue loop (descriptors in its own registers), the other lanes wait at the end
        if (elect_one()) {
          while (done_warps < nwarps) {
#pragma unroll 1
            for (int x = 0; x < 2; ++x) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
                umma_ss<MmaKind::F16, 1, false>(tm + x * 128, make_desc_kmajor_sw128(q_smem + x * 32768 + off), make_desc_kmajor_sw128(k_smem + off), idesc_qk, ks != 0);
              }
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
                if (mma_mode == 1)
                  umma_ss<MmaKind::F16, 1, false>(tm + 256 + x * 128, make_desc_kmajor_sw128(base + x * 32768 + off), make_desc_mnmajor_sw128(v_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
                else
                  umma_ts<MmaKind::F16, false>(tm + 256 + x * 128, tm + x * 128 + ks * 8, make_desc_mnmajor_sw128(v_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
              }
            }
            n_mma += 32;
          }
        }
        __syncwarp();
        int mx = 0;
        for (int l = 0; l < 32; ++l) mx = max(mx, __shfl_sync(0xffffffffu, (int)n_mma, l));
        n_mma = mx;
      } else
      while (done_warps < nwarps) {
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
            umma_ss<MmaKind::F16, 1, true>(tm + x * 128, make_desc_kmajor_sw128(q_smem + x * 32768 + off), make_desc_kmajor_sw128(k_smem + off), idesc_qk, ks != 0);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
            if (mma_mode == 1)
              umma_ss<MmaKind::F16, 1, true>(tm + 256 + x * 128, make_desc_kmajor_sw128(base + x * 32768 + off), make_desc_mnmajor_sw128(v_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
            else
              umma_ts<MmaKind::F16, true>(tm + 256 + x * 128, tm + x * 128 + ks * 8, make_desc_mnmajor_sw128(v_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
          }
        }
        n_mma += 32;
      }
      tc_commit_elect(smem_u32(&mbar));
      mbar_wait(smem_u32(&mbar), 0);
      const long long t1 = clock64();
      if (lane == 0) { out[148 * 8 * 5 + blockIdx.x * 2] = t1 - t0; out[148 * 8 * 5 + blockIdx.x * 2 + 1] = n_mma; }
    }
  } else {
    reg_alloc<208>();
    const int x = warp >> 2, lg = warp & 3;
    const int row_in_tile = lg * 32 + lane;
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    const uint32_t tS = tm + lane_off + (uint32_t)(x * 128);
    const uint32_t p_row = base + (uint32_t)x * 32768u + (uint32_t)row_in_tile * 128u;
    const uint32_t p_sw = (uint32_t)(row_in_tile & 7);
    {  // S <- deterministic values in [-24, 24] (scale_log2 ~ 0.1275: exponents within [-6, 0] of the max)
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint((float)(((row_in_tile * 131 + (c * 32 + i) * 71) % 97) - 48) * 0.5f);
        tmem_st_32x32(tS + c * 32, v);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;");
    tc_fence_after();
    if (warp < nwarps) {
      float m_run = 24.f * scale_log2 - 1.f, l_run = 0.f;   // (a running max as after the first tiles: no rescale in the loop)
      long long t_ld = 0, t_max = 0, t_exp = 0, t_st = 0;
      const long long t0 = clock64();
      for (int t = 0; t < tiles; ++t) {
        const uint64_t scale2 = f2(scale_log2, scale_log2);
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
        if constexpr (VAR == 0) {
          long long c0 = ST ? clock64() : 0;
          uint32_t s0[32], s1[32], s2[32], s3[32];
          tmem_ld_32x32(tS, s0); tmem_ld_32x32(tS + 32u, s1); tmem_ld_32x32(tS + 64u, s2); tmem_ld_32x32(tS + 96u, s3);
          tmem_ld_wait();
          long long c1 = ST ? clock64() : 0;
          const float mx = fmaxf(fmax3(max32(s0), max32(s1), max32(s2)), max32(s3));
          const float m_new = fmaxf(m_run, mx * scale_log2);
          if (__any_sync(0xffffffffu, m_new > m_run + 8.f)) m_run = m_new;
          long long c2 = ST ? clock64() : 0;
          const uint64_t negm2 = f2(-m_run, -m_run);
          uint32_t pk0[16], pk1[16], pk2[16], pk3[16];
          compute32<EMU>(s0, pk0, scale2, negm2, acc); compute32<EMU>(s1, pk1, scale2, negm2, acc);
          compute32<EMU>(s2, pk2, scale2, negm2, acc); compute32<EMU>(s3, pk3, scale2, negm2, acc);
          long long c3 = ST ? clock64() : 0;
          store32(p_row, p_sw, pk0, 0); store32(p_row, p_sw, pk1, 1); store32(p_row, p_sw, pk2, 2); store32(p_row, p_sw, pk3, 3);
          fence_proxy_async_smem();
          long long c4 = ST ? clock64() : 0;
          t_ld += c1 - c0; t_max += c2 - c1; t_exp += c3 - c2; t_st += c4 - c3;
        } else if constexpr (VAR == 1) {
          long long c0 = ST ? clock64() : 0;
          uint32_t s0[32], s1[32], s2[32], s3[32];
          tmem_ld_32x32(tS, s0); tmem_ld_32x32(tS + 32u, s1); tmem_ld_32x32(tS + 64u, s2); tmem_ld_32x32(tS + 96u, s3);
          tmem_ld_wait();
          long long c1 = ST ? clock64() : 0;
          const uint64_t negm2 = f2(-m_run, -m_run);
          uint32_t pk0[16], pk1[16], pk2[16], pk3[16];
          compute32<EMU>(s0, pk0, scale2, negm2, acc); const float x0 = max32(s0);
          compute32<EMU>(s1, pk1, scale2, negm2, acc); const float x1 = max32(s1);
          compute32<EMU>(s2, pk2, scale2, negm2, acc); const float x2 = max32(s2);
          compute32<EMU>(s3, pk3, scale2, negm2, acc); const float x3 = max32(s3);
          const float mx = fmaxf(fmax3(x0, x1, x2), x3);
          long long c3 = ST ? clock64() : 0;
          store32(p_row, p_sw, pk0, 0); store32(p_row, p_sw, pk1, 1); store32(p_row, p_sw, pk2, 2); store32(p_row, p_sw, pk3, 3);
          fence_proxy_async_smem();
          const float m_new = fmaxf(m_run, mx * scale_log2);
          if (__any_sync(0xffffffffu, m_new > m_run + 8.f)) { m_run = m_new; l_run = 0.f; }   // (miss: the kernel would redo the tile)
          long long c4 = ST ? clock64() : 0;
          t_ld += c1 - c0; t_exp += c3 - c1; t_st += c4 - c3;
        } else if constexpr (VAR == 2) {
          long long c0 = ST ? clock64() : 0;
          uint32_t s0[32], s1[32], s2[32], s3[32];
          tmem_ld_32x32(tS, s0);
          tmem_ld_wait();
          tmem_ld_32x32(tS + 32u, s1); tmem_ld_32x32(tS + 64u, s2); tmem_ld_32x32(tS + 96u, s3);
          long long c1 = ST ? clock64() : 0;
          const uint64_t negm2 = f2(-m_run, -m_run);
          uint32_t pk0[16], pk1[16], pk2[16], pk3[16];
          compute32<EMU>(s0, pk0, scale2, negm2, acc); const float x0 = max32(s0);
          tmem_ld_wait();
          compute32<EMU>(s1, pk1, scale2, negm2, acc); const float x1 = max32(s1);
          compute32<EMU>(s2, pk2, scale2, negm2, acc); const float x2 = max32(s2);
          compute32<EMU>(s3, pk3, scale2, negm2, acc); const float x3 = max32(s3);
          const float mx = fmaxf(fmax3(x0, x1, x2), x3);
          long long c3 = ST ? clock64() : 0;
          store32(p_row, p_sw, pk0, 0); store32(p_row, p_sw, pk1, 1); store32(p_row, p_sw, pk2, 2); store32(p_row, p_sw, pk3, 3);
          fence_proxy_async_smem();
          const float m_new = fmaxf(m_run, mx * scale_log2);
          if (__any_sync(0xffffffffu, m_new > m_run + 8.f)) { m_run = m_new; l_run = 0.f; }
          long long c4 = ST ? clock64() : 0;
          t_ld += c1 - c0; t_exp += c3 - c1; t_st += c4 - c3;
        } else if constexpr (VAR == 3) {
          // probes (EMU selects): 0 = 128 MUFU, dependent only on the loop-carried value; 1 = 256 FFMA2 (8 chains);
          // 2 = 128 MUFU each followed by 2 FFMA2; 3 = 128 MUFU + 4 FFMA2 each; 4 = 128 MUFU + 4 scalar FFMA each;
          // 5 = 128 F2FP; 6 = 128 MUFU + 1 F2FP per 2
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = m_run + (float)i;
          uint64_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = f2(l_run + (float)i, m_run);
          long long c0 = ST ? clock64() : 0;
#pragma unroll
          for (int i = 0; i < (EMU >= 7 ? 0 : 128); ++i) {
            if (EMU == 0 || EMU == 2 || EMU == 3 || EMU == 4 || EMU == 6) v[i & 7] = ex2(v[i & 7]);
            if (EMU == 1) { w[i & 7] = ffma2(w[i & 7], scale2, w[(i + 1) & 7]); w[(i + 3) & 7] = ffma2(w[(i + 3) & 7], scale2, w[(i + 5) & 7]); }
            if (EMU == 2) { w[i & 7] = ffma2(w[i & 7], scale2, scale2); w[(i + 3) & 7] = ffma2(w[(i + 3) & 7], scale2, scale2); }
            if (EMU == 3) {
#pragma unroll
              for (int k = 0; k < 4; ++k) w[(i + 2 * k) & 7] = ffma2(w[(i + 2 * k) & 7], scale2, scale2);
            }
            if (EMU == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) { float a, b; unf2(w[(i + 2 * k) & 7], a, b); a = fmaf(a, scale_log2, 0.5f); w[(i + 2 * k) & 7] = f2(a, b); }
            }
            if (EMU == 5) { float a, b; unf2(w[i & 7], a, b); const uint32_t pkd = pk_bf16(a, b); w[i & 7] = f2(__uint_as_float(pkd), b); }
            if (EMU == 6 && (i & 1)) { const uint32_t pkd = pk_bf16(v[i & 7], v[(i - 1) & 7]); w[i & 7] = f2(__uint_as_float(pkd), 0.f); }
          }
          if (EMU == 9) __nanosleep(1500);
          if (EMU == 7) {
            uint32_t r0[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) { tmem_ld_32x32(tS + c * 32, r0); tmem_ld_wait(); v[c] += __uint_as_float(r0[c]); }
          }
          if (EMU == 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t pkz[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) pkz[i] = __float_as_uint(v[i & 7]) + i;
              store32(p_row, p_sw, pkz, c);
            }
            fence_proxy_async_smem();
          }
          long long c1 = ST ? clock64() : 0;
          t_exp += c1 - c0;
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { float p, q; unf2(w[i], p, q); a += v[i] + p + q; }
          m_run = a * 1e-30f + 1.f;
          l_run = a * 1e-30f;
        }
        if constexpr (VAR != 3) {
          float a0, a1, b0, b1;
          unf2(fadd2(acc[0], acc[1]), a0, a1);
          unf2(fadd2(acc[2], acc[3]), b0, b1);
          l_run += (a0 + a1) + (b0 + b1);
        }
      }
      const long long t1 = clock64();
      if (lane == 0) {
        long long* o = out + ((size_t)blockIdx.x * 8 + warp) * 5;
        o[0] = t1 - t0; o[1] = t_ld; o[2] = t_max; o[3] = t_exp; o[4] = t_st;
      }
      if (l_run == 12345.678f) sink[0] = l_run + m_run;
      __syncwarp();
      if (lane == 0) atomicAdd((int*)&done_warps, 1);
    }
  }
  __syncthreads();
  if (warp == 8) tmem_dealloc<1>(tm, 512);
}

template <int VAR, int EMU, bool ST = true>
static void run(const char* name, long long* d, float* sink, int mma_mode = 0) {
  const int smem = 200 * 1024, tiles = 400;
  cudaFuncSetAttribute(lab<VAR, EMU, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int nw : {4, 8}) {
    cudaMemset(d, 0, 8 * (148 * 8 * 5 + 296));
    lab<VAR, EMU, ST><<<148, 384, smem>>>(nw, tiles, 0.1275f, d, sink, mma_mode);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 8 * 5 + 296];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double tot = 0, ld = 0, mx = 0, ex = 0, st = 0;
    int n = 0;
    for (int b = 0; b < 148; ++b)
      for (int w = 0; w < nw; ++w) {
        const long long* o = h + ((size_t)b * 8 + w) * 5;
        tot += o[0]; ld += o[1]; mx += o[2]; ex += o[3]; st += o[4];
        ++n;
      }
    const double k = 1.0 / ((double)n * tiles);
    printf("%-44s %d warp(s)/SMSP: %7.1f cycles/tile/warp (ld %6.1f max %6.1f exp %6.1f st %6.1f) -> SMSP period for 2 tiles %7.1f  %s\n", name,
           nw / 4, tot * k, ld * k, mx * k, ex * k, st * k, nw == 8 ? tot * k : 2 * tot * k, e == cudaSuccess ? "" : cudaGetErrorString(e));
    {
      double per[4] = {0, 0, 0, 0};
      for (int b = 0; b < 148; ++b)
        for (int w = 0; w < nw; ++w) per[w & 3] += (double)h[((size_t)b * 8 + w) * 5] / (148.0 * (nw / 4) * tiles);
      printf("      per scheduler (warp %% 4): %7.1f %7.1f %7.1f %7.1f\n", per[0], per[1], per[2], per[3]);
    }
    if (mma_mode) printf("      with the MMA stream (%s): %.1f cycles / MMA\n", mma_mode == 1 ? "P from smem" : "P from TMEM", (double)h[148 * 8 * 5] / (double)h[148 * 8 * 5 + 1]);
  }
}

int main(int argc, char**) {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  long long* d; cudaMalloc(&d, 8 * (148 * 8 * 5 + 296)); float* sink; cudaMalloc(&sink, 4);
  const bool probes = argc > 1;
  if (probes) {
    run<3, 0>("probe: 128 MUFU.EX2", d, sink);
    run<3, 1>("probe: 256 FFMA2", d, sink);
    run<3, 3>("probe: 128 x (MUFU + 4 FFMA2)", d, sink);
    run<3, 5>("probe: 128 F2FP", d, sink);
    run<0, 4>("v0 phases, EMU 4 (kernel today)", d, sink);
    run<2, 8>("v2 speculative max + chunked ld, EMU 8", d, sink);
  }
  // does the issuing warp (warp 9, scheduler 1) slow the softmax warps it shares a scheduler with?
  for (int mm : {0, 3, 5, 6}) {
    printf("---- warp 9: %s\n", mm == 0 ? "idle" : mm == 3 ? "one thread spinning on mbarrier.try_wait" : mm == 5 ? "one thread issuing MMAs flat out (SS PV), blocked on the full queue most of the time" : "one thread issuing MMAs flat out (TS PV)");
    run<0, 4, false>("softmax v0 EMU 4", d, sink, mm);
  }
  return 0;
}
