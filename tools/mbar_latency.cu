// Micro-benchmark: cost of waiting on an ALREADY satisfied mbarrier phase (try_wait vs test_wait),
// of tcgen05.fence::after_thread_sync, and of a trace stamp, as seen by a single thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -I include -o build/mbar_latency tools/mbar_latency.cu
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

__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void bench(long long* out, long long* scratch) {
  __shared__ uint64_t bars[64];
  const uint32_t b0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(b0 + 8 * i, 1);
    fence_mbar_init();
    for (int i = 0; i < 64; ++i) mbar_arrive(b0 + 8 * i);  // phase 0 of every barrier is complete
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0, t1;
    const int N = 32;
    t0 = clock64();
    for (int i = 0; i < N; ++i) mbar_wait(b0 + 8 * i, 0);
    t1 = clock64();
    out[0] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) while (!mbar_test_wait(b0 + 8 * (32 + i), 0)) {}
    t1 = clock64();
    out[1] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { mbar_wait(b0 + 8 * i, 0); tc_fence_after(); }
    t1 = clock64();
    out[2] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { while (!mbar_test_wait(b0 + 8 * i, 0)) {} tc_fence_after(); }
    t1 = clock64();
    out[3] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) scratch[i] = clock64();
    t1 = clock64();
    out[4] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) tc_fence_after();
    t1 = clock64();
    out[5] = (t1 - t0) / N;
  }
}

int main() {
  long long *d, *s; cudaMalloc(&d, 64); cudaMalloc(&s, 8 * 64);
  bench<<<1, 32>>>(d, s);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("%s\n satisfied try_wait %lld cyc | test_wait %lld | try_wait+fence %lld | test_wait+fence %lld | clock64+store %lld | fence %lld\n",
         cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4], h[5]);
  return 0;
}
