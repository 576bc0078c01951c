// Micro-benchmark: sustained rate of tcgen05.mma.cta_group::2 (M=256, N=128, K=16, bf16, SS) for the two operand
// layouts of the attention kernel, against the single-CTA M=128 N=128 form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -I include -o build/cg2_rate tools/cg2_rate.cu
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

template <int CG>
__global__ void __launch_bounds__(128, 1) rate(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const uint32_t bar = base + 131072;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 131072 + 32);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<CG>(smem_u32(tmem_ptr), 512);
  tc_fence_before(); __syncthreads();
  if (CG == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tm = *tmem_ptr;
  if (mode == 3) {
    // two issuing threads in different warps, each streaming its own 8-MMA groups
    const uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128 * CG, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128 * CG, 128, 0, 1);
    const uint32_t a_smem = base, p_smem = base + 32768, k_smem = base + 65536, v_smem = base + 98304;
    const uint32_t bar2 = bar + 8;
    if (rank == 0 && (threadIdx.x == 0 || threadIdx.x == 32)) {
      const bool second = threadIdx.x == 32;
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
          const uint32_t koff = (uint32_t)(ks / 4) * (16384u / CG) + (uint32_t)(ks % 4) * 32u;
          if (!second) umma_ss<MmaKind::F16, CG>(tm, make_desc_kmajor_sw128(a_smem + off), make_desc_kmajor_sw128(k_smem + koff), idesc_qk, ks != 0);
          else umma_ss<MmaKind::F16, CG>(tm + 256, make_desc_kmajor_sw128(p_smem + off), make_desc_mnmajor_sw128(v_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
        }
      }
      const uint32_t b = second ? bar2 : bar;
      if (CG == 2) tc_commit_cg2(b, 0b11); else tc_commit(b);
      mbar_wait(b, 0);
      long long t1 = clock64();
      out[blockIdx.x * 2 + (second ? 1 : 0)] = t1 - t0;
    } else if (threadIdx.x == 0 || threadIdx.x == 32) {
      mbar_wait(threadIdx.x == 32 ? bar2 : bar, 0);
    }
  } else if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128 * CG, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 128 * CG, 128, 0, 1);
    const uint32_t a_smem = base, p_smem = base + 32768, k_smem = base + 65536, v_smem = base + 98304;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode != 1) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
          const uint32_t koff = (uint32_t)(ks / 4) * (16384u / CG) + (uint32_t)(ks % 4) * 32u;
          umma_ss<MmaKind::F16, CG>(tm, make_desc_kmajor_sw128(a_smem + off), make_desc_kmajor_sw128(k_smem + koff), idesc_qk, ks != 0);
        }
      }
      if (mode != 0) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (uint32_t)(ks / 4) * 16384u + (uint32_t)(ks % 4) * 32u;
          umma_ss<MmaKind::F16, CG>(tm + 256, make_desc_kmajor_sw128(p_smem + off), make_desc_mnmajor_sw128(v_smem + ks * 2048u, 16384, 1024), idesc_pv, 1);
        }
      }
    }
    if (CG == 2) tc_commit_cg2(bar, 0b11); else tc_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (threadIdx.x == 0) {
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (threadIdx.x < 32) tmem_dealloc<CG>(tm, 512);
}

template <int CG>
void run(const char* name, int mode) {
  long long* d; cudaMalloc(&d, 8 * 296); cudaMemset(d, 0, 8 * 296);
  const int smem = 140 * 1024, iters = 2000;
  cudaFuncSetAttribute(rate<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, rate<CG>, mode, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, 8 * 148, cudaMemcpyDeviceToHost);
  const int per = mode >= 2 ? 16 : 8;
  printf("%-44s %6.1f cycles / MMA  %s\n", name, (double)h[0] / (iters * (double)per), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<1>("cta_group::1 M128 N128  QK-like (K-major B)", 0);
  run<1>("cta_group::1 M128 N128  PV-like SS (MN-major B)", 1);
  run<1>("cta_group::1 M128 N128  QK + PV", 2);
  run<2>("cta_group::2 M256 N128  QK-like (K-major B)", 0);
  run<2>("cta_group::2 M256 N128  PV-like SS (MN-major B)", 1);
  run<2>("cta_group::2 M256 N128  QK + PV", 2);
  run<1>("cta_group::1 two issuing warps (QK | PV)", 3);
  run<2>("cta_group::2 two issuing warps (QK | PV)", 3);
  return 0;
}
