// Micro-benchmark: (1) how many tcgen05.mma a thread can issue before the issue itself blocks (queue depth),
// (2) whether the sustained MMA rate on all SMs drops below 64 cycles/MMA when run long enough to hit the power cap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -I include -o build/mma_queue tools/mma_queue.cu
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

__global__ void __launch_bounds__(128, 1) queue_depth(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  if (threadIdx.x < 32) tmem_alloc<1>(smem_u32(&tmem_ptr), 512);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 0);
    const uint64_t a = make_desc_kmajor_sw128(base), b = make_desc_kmajor_sw128(base + 32768);
    uint32_t phase = 0;
    for (int n = 1; n <= 32; ++n) {
      long long t0 = clock64();
      for (int i = 0; i < n; ++i) umma_ss<MmaKind::F16, 1>(tm, a, b, idesc, 1);
      long long t1 = clock64();
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      long long t2 = clock64();
      out[2 * n] = t1 - t0;
      out[2 * n + 1] = t2 - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tm, 512);
}

__global__ void __launch_bounds__(128, 1) sustained(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  if (threadIdx.x < 32) tmem_alloc<1>(smem_u32(&tmem_ptr), 512);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 0);
    const uint64_t a = make_desc_kmajor_sw128(base), b = make_desc_kmajor_sw128(base + 32768);
    // report cycles/MMA for 8 consecutive windows
    for (int w = 0; w < 8; ++w) {
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma_ss<MmaKind::F16, 1>(tm + (i & 1) * 128, a, b, idesc, 1);
      }
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), w & 1);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[w] = t1 - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 128); cudaMemset(d, 0, 8 * 128);
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(queue_depth, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(sustained, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  queue_depth<<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[128]; cudaMemcpy(h, d, 8 * 128, cudaMemcpyDeviceToHost);
  printf("queue depth probe (%s): n MMAs -> cycles until issue returns / until complete\n", cudaGetErrorString(e));
  for (int n = 1; n <= 32; ++n) printf("  n=%2d issue %5lld  done %5lld\n", n, h[2 * n], h[2 * n + 1]);
  // sustained: 8 windows x iters x 16 MMAs on all SMs; each window ~ iters*16*64 cycles
  const int iters = 40000;  // 41 M cycles ~ 25 ms per window, 0.2 s in all
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  sustained<<<148, 128, smem>>>(iters, d);
  cudaEventRecord(e1);
  e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(h, d, 8 * 8, cudaMemcpyDeviceToHost);
  printf("sustained on 148 SMs (%s), %.1f ms total:\n", cudaGetErrorString(e), ms);
  for (int w = 0; w < 8; ++w) printf("  window %d: %.2f cycles / MMA\n", w, (double)h[w] / (iters * 16.0));
  printf("  => %.0f TFLOP/s dense bf16, SM clock ~%.0f MHz\n", 148.0 * 8 * iters * 16 * 2.0 * 128 * 128 * 16 / (ms * 1e-3) / 1e12,
         (double)(h[0] + h[1] + h[2] + h[3] + h[4] + h[5] + h[6] + h[7]) / (ms * 1e-3) / 1e6);
  return 0;
}
