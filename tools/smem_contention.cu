// Micro-benchmark: does generic shared-memory store traffic slow down tcgen05.mma operand fetch?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -o /tmp/smem_contention tools/smem_contention.cu
// One CTA per SM: thread 0 streams M128 N128 K16 bf16 MMAs (PV then QK, attention-like; PV either TS or SS),
// warps 4..11 store 16 B / thread / iteration into a scratch region with `delay` dependent FMAs between stores.
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

template <int PV_SS>
__global__ void __launch_bounds__(384, 1) bench(int iters, int delay, int stores_on, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  if (threadIdx.x < 32) tmem_alloc<1>(smem_u32(&tmem_ptr), 512);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); stop = 0; }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const uint32_t a_smem = base, b_smem = base + 32768, p_smem = base + 98304, scratch = base + 131072;
  if (threadIdx.x == 0) {
    const uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128, 128, 0, 1);
    const uint32_t tO = tm + 256, tS = tm;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t vdesc = make_desc_mnmajor_sw128(b_smem + ks * 2048u, 16384, 1024);
        if (PV_SS) {
          const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
          umma_ss<MmaKind::F16, 1>(tO, make_desc_kmajor_sw128(p_smem + off), vdesc, idesc_pv, 1);
        } else {
          umma_ts<MmaKind::F16>(tO, tm + 128 + ks * 8, vdesc, idesc_pv, 1);
        }
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
        umma_ss<MmaKind::F16, 1>(tS, make_desc_kmajor_sw128(a_smem + off), make_desc_kmajor_sw128(b_smem + 32768 + off), idesc_qk, ks != 0);
      }
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    stop = 1;
  } else if (threadIdx.x >= 128 && stores_on) {
    // conflict-free 16-byte stores: thread i of the 256 writes chunk i of a 4 KB line, lines rotate over 64 KB
    const uint32_t tid = threadIdx.x - 128;
    long long n = 0;
    float f = (float)tid;
    while (!stop) {
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        for (int d = 0; d < delay; ++d) f = fmaf(f, 1.0001f, 0.5f);
        const uint32_t addr = scratch + (uint32_t)(r * 4096) + tid * 16u;
        const uint32_t v = __float_as_uint(f);
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(v) : "memory");
      }
      n += 16;
    }
    if (tid == 0) out[blockIdx.x * 2 + 1] = n * 256 * 16;  // bytes stored by the CTA
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tm, 512);
}

template <int PV_SS>
void run(const char* name, int delay, int stores_on) {
  const int grid = 148;
  long long* d; cudaMalloc(&d, 16 * grid); cudaMemset(d, 0, 16 * grid);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench<PV_SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  bench<PV_SS><<<grid, 384, smem>>>(iters, delay, stores_on, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-34s delay=%3d: %6.1f cycles / MMA, generic stores %6.1f B/clk   %s\n", name, delay, (double)h[0] / (iters * 16.0),
         (double)h[1] / (double)h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<0>("PV TS + QK SS, no stores", 0, 0);
  run<1>("PV SS + QK SS, no stores", 0, 0);
  for (int delay : {0, 4, 16, 64}) {
    run<0>("PV TS + QK SS, 8 warps storing", delay, 1);
    run<1>("PV SS + QK SS, 8 warps storing", delay, 1);
  }
  return 0;
}
