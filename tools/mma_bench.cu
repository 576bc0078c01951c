// Micro-benchmark: issue rate of tcgen05.mma for the shapes the attention kernel uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -o /tmp/mma_bench tools/mma_bench.cu
// Prints SM cycles per MMA for: SS (A,B from smem) vs TS (A from TMEM), N = 64/128/256, bf16 and fp8.
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

template <int MODE /*0 SS bf16, 1 TS bf16, 2 SS fp8, 3 SS bf16 B MN-major, 4 TS bf16 B MN-major, 5/6 attention-like*/>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  if (threadIdx.x < 32) tmem_alloc<1>(smem_u32(&tmem_ptr), 512);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t a_smem = base, b_smem = base + 65536;
    uint32_t idesc;
    if (MODE == 2) idesc = make_idesc(kFmtE4M3, kFmtE4M3, kAccF32, 128, N, 0, 0);
    else idesc = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, N, 0, (MODE >= 3) ? 1 : 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
        uint64_t bdesc = (MODE >= 3) ? make_desc_mnmajor_sw128(b_smem + ks * 2048u, 16384, 1024)
                                     : make_desc_kmajor_sw128(b_smem + off);
        if (MODE == 0 || MODE == 3) umma_ss<MmaKind::F16, 1>(tm, make_desc_kmajor_sw128(a_smem + off), bdesc, idesc, 1);
        else if (MODE == 2) umma_ss<MmaKind::F8F6F4, 1>(tm, make_desc_kmajor_sw128(a_smem + off), bdesc, idesc, 1);
        else umma_ts<MmaKind::F16>(tm, tm + 256 + ks * 8, bdesc, idesc, 1);
      }
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
    if (MODE >= 5) {
      // attention-like stream: PV (TS: A = P read from TMEM region X, D = O) followed by QK (SS, D = S).
      // MODE 5: S aliases X (as in the kernel: P lives over S), MODE 6: S in a different region.
      const uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 1);
      const uint32_t tX = tm, tO = tm + 256, tS = (MODE == 5) ? tm : tm + 128;
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ts<MmaKind::F16>(tO, tX + ks * 8, make_desc_mnmajor_sw128(b_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
          umma_ss<MmaKind::F16, 1>(tS, make_desc_kmajor_sw128(a_smem + off), make_desc_kmajor_sw128(b_smem + 32768 + off), idesc_qk, ks != 0);
        }
      }
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 1);
      t1 = clock64();
      out[0] = (t1 - t0) / 2;  // 16 MMAs per iteration, caller divides by 8
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tm, 512);
}

template <int MODE>
void run(const char* name, int N, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  bench<MODE><<<grid, 128, smem>>>(N, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d grid=%3d: %7.1f cycles / MMA (K=%d)   %s\n", name, N, grid, (double)h / (iters * 8.0), MODE == 2 ? 32 : 16,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    for (int N : {64, 128, 256}) {
      run<0>("SS bf16 (A,B K-major smem)", N, grid);
      run<1>("TS bf16 (A tmem, B K-major)", N, grid);
      run<3>("SS bf16 (B MN-major)", N, grid);
      run<4>("TS bf16 (A tmem, B MN-major)", N, grid);
      run<2>("SS fp8  (K=32)", N, grid);
    }
    run<5>("PV(TS)+QK(SS), S aliases P", 128, grid);
    run<6>("PV(TS)+QK(SS), S elsewhere", 128, grid);
  }
  return 0;
}
