// Micro-benchmark: what slows tcgen05.mma below its 64-cycle floor inside the attention kernel?
// Thread 0 of warp 4 streams M128 N128 K16 bf16 MMAs (PV(TS)+QK(SS) pattern); warps 0-3 / 5-8 optionally run one of:
//   mode 1: tcgen05.ld 32x32b.x32 loops (4 x 32 columns = an S row) from columns the MMAs do not write
//   mode 2: tcgen05.ld of the S columns the MMAs are writing (real kernel: the other tile's S)
//   mode 3: tcgen05.st 32x32b.x16 loops
//   mode 4: MUFU ex2 + FMA loops (issue-slot / SMSP contention only)
//   mode 5: cp.async.bulk global->shared 16 KB copies in flight (TMA-like smem writes)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -I include -o build/mma_contention tools/mma_contention.cu
#include <cstdio>
#include "sm100.cuh"
namespace fdm {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -3; }
int require_sm100() { return 0; }
int num_sms() { return 148; }
}
using namespace fdm;
using namespace fdm::sm100;

__global__ void __launch_bounds__(384, 1) bench(int mode, int iters, const uint8_t* gsrc, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar, cbar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4) tmem_alloc<1>(smem_u32(&tmem_ptr), 512);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&cbar), 1); fence_mbar_init(); stop = 0; }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 1);
      const uint32_t a_smem = base, b_smem = base + 32768;
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ts<MmaKind::F16>(tm + 256, tm + 128 + ks * 8, make_desc_mnmajor_sw128(b_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
          umma_ss<MmaKind::F16, 1>(tm, make_desc_kmajor_sw128(a_smem + off), make_desc_kmajor_sw128(b_smem + 32768 + off), idesc_qk, ks != 0);
        }
      }
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (warp < 4 || (warp >= 5 && warp < 9)) {
    const int lg = warp < 4 ? warp : warp - 5;
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    float acc = 0.f;
    uint32_t r[32];
    if (mode == 1 || mode == 2) {
      const uint32_t col = mode == 1 ? 384u : 0u;
      while (!stop) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x32(tm + lane_off + col + c * 32, r);
          tmem_ld_wait();
          acc += __uint_as_float(r[lane & 31]);
        }
      }
    } else if (mode == 3) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = i;
      while (!stop) {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st_32x16(tm + lane_off + 384u + c * 16, v);
        tmem_st_wait();
      }
    } else if (mode == 4) {
      float x = (float)lane * 0.01f;
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          x = fmaf(y, 0.001f, x * 0.5f);
        }
      }
      acc = x;
    } else if (mode == 5) {
      if (warp == 0 && lane == 0) {
        uint32_t ph = 0;
        while (!stop) {
          mbar_arrive_expect_tx(smem_u32(&cbar), 4 * 16384);
          for (int i = 0; i < 4; ++i)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             base + 98304 + i * 16384),
                         "l"(gsrc + (size_t)blockIdx.x * 65536 + i * 16384), "r"(16384), "r"(smem_u32(&cbar))
                         : "memory");
          mbar_wait(smem_u32(&cbar), ph);
          ph ^= 1;
          acc += 1.f;
        }
        out[148 + blockIdx.x] = (long long)acc * 65536;
      }
    }
    if (acc == 12345.678f) sink[0] = acc;
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc<1>(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 296); float* sink; cudaMalloc(&sink, 4);
  uint8_t* g; cudaMalloc(&g, 148 * 65536); cudaMemset(g, 1, 148 * 65536);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"MMA alone", "+ tcgen05.ld (idle columns)", "+ tcgen05.ld (columns being written)", "+ tcgen05.st", "+ MUFU/FMA loops", "+ bulk copies into smem"};
  const int iters = 2000;
  for (int mode = 0; mode <= 5; ++mode) {
    cudaMemset(d, 0, 8 * 296);
    bench<<<148, 384, smem>>>(mode, iters, g, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, d, 8 * 296, cudaMemcpyDeviceToHost);
    printf("%-40s %6.1f cycles / MMA", names[mode], (double)h[0] / (iters * 16.0));
    if (mode == 5) printf("   (bulk copy %.1f B/clk)", (double)h[148] / (double)h[0]);
    printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
