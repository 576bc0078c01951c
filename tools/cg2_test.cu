// Functional check of the cta_group::2 building blocks used by the 2-CTA attention kernel:
// cluster of 2 CTAs, tcgen05.alloc.cta_group::2, one M=256 N=128 tcgen05.mma.cta_group::2 (K-major B = "QK"
// and MN-major B = "PV"), multicast commit, remote mbarrier arrive. Compares with a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdm_b200/csrc -I include -o build/cg2_test tools/cg2_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <vector>
#include "sm100.cuh"
namespace fdm {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -3; }
int require_sm100() { return 0; }
int num_sms() { return 148; }
}
using namespace fdm;
using namespace fdm::sm100;

constexpr int KD = 64;  // reduction extent (one 128-byte swizzled row of bf16)

// element (row, col) of a [rows x 64] bf16 tile stored as 128-byte rows with the 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return (uint32_t)(row * 128 + ((((col * 2) >> 4) ^ (row & 7)) << 4) + ((col * 2) & 15));
}

// A: [256 x KD] (CTA r owns rows 128r..), Bk: [128 n x KD] K-major (CTA r owns n = 64r..), Bm: [KD k x 128 n]
// MN-major (CTA r owns n columns 64r..). out_qk / out_pv: [256 x 128] fp32.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cg2_kernel(const __nv_bfloat16* A, const __nv_bfloat16* Bk, const __nv_bfloat16* Bm, float* out_qk, float* out_pv) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_off = 0, bk_off = 16384, bm_off = 32768, bar_off = 49152;
  const uint32_t rank = cluster_ctarank();
  const uint32_t done_bar = base + bar_off, ready_bar = base + bar_off + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + bar_off + 16);
  if (threadIdx.x == 0) {
    mbar_init(done_bar, 1);
    mbar_init(ready_bar, 2);  // one elected arrival per CTA
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc<2>(smem_u32(tmem_ptr), 256);
  // fill operands (generic stores)
  for (int i = threadIdx.x; i < 128 * KD; i += 128) {
    const int r = i / KD, c = i % KD;
    *reinterpret_cast<__nv_bfloat16*>(smem + a_off + sw128_off(r, c)) = A[(rank * 128 + r) * KD + c];
  }
  for (int i = threadIdx.x; i < 64 * KD; i += 128) {
    const int n = i / KD, c = i % KD;
    *reinterpret_cast<__nv_bfloat16*>(smem + bk_off + sw128_off(n, c)) = Bk[(rank * 64 + n) * KD + c];
  }
  for (int i = threadIdx.x; i < KD * 64; i += 128) {
    const int k = i / 64, n = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(smem + bm_off + sw128_off(k, n)) = Bm[k * 128 + rank * 64 + n];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = *tmem_ptr;
  // every CTA tells the leader its operands are in place (remote arrive for rank 1)
  if (threadIdx.x == 0) mbar_arrive_cluster(ready_bar, 0);
  if (rank == 0 && threadIdx.x == 0) {
    mbar_wait(ready_bar, 0);
    tc_fence_after();
    const uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 256, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kAccF32, 256, 128, 0, 1);
    for (int ks = 0; ks < KD / 16; ++ks) {
      umma_ss<MmaKind::F16, 2>(tm, make_desc_kmajor_sw128(base + a_off + ks * 32), make_desc_kmajor_sw128(base + bk_off + ks * 32),
                               idesc_qk, ks != 0);
    }
    for (int ks = 0; ks < KD / 16; ++ks) {
      umma_ss<MmaKind::F16, 2>(tm + 128, make_desc_kmajor_sw128(base + a_off + ks * 32),
                               make_desc_mnmajor_sw128(base + bm_off + ks * 16 * 128, 16384, 1024), idesc_pv, ks != 0);
    }
    tc_commit_cg2(done_bar, 0b11);
  }
  mbar_wait(done_bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const int row = rank * 128 + warp * 32 + lane;
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(tm + lane_off + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out_qk[row * 128 + c * 32 + i] = __uint_as_float(r[i]);
    tmem_ld_32x32(tm + 128 + lane_off + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out_pv[row * 128 + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) tmem_dealloc<2>(tm, 256);
}

int main() {
  std::vector<__nv_bfloat16> A(256 * KD), Bk(128 * KD), Bm(KD * 128);
  std::vector<float> Af(256 * KD), Bkf(128 * KD), Bmf(KD * 128);
  srand(1);
  auto rnd = [] { return (float)((rand() % 9) - 4); };
  for (size_t i = 0; i < A.size(); ++i) { Af[i] = rnd(); A[i] = __float2bfloat16(Af[i]); }
  for (size_t i = 0; i < Bk.size(); ++i) { Bkf[i] = rnd(); Bk[i] = __float2bfloat16(Bkf[i]); }
  for (size_t i = 0; i < Bm.size(); ++i) { Bmf[i] = rnd(); Bm[i] = __float2bfloat16(Bmf[i]); }
  __nv_bfloat16 *dA, *dBk, *dBm; float *dq, *dp;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dBk, Bk.size() * 2); cudaMalloc(&dBm, Bm.size() * 2);
  cudaMalloc(&dq, 256 * 128 * 4); cudaMalloc(&dp, 256 * 128 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBk, Bk.data(), Bk.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBm, Bm.data(), Bm.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dq, 0xff, 256 * 128 * 4); cudaMemset(dp, 0xff, 256 * 128 * 4);
  const int smem = 52 * 1024;
  cudaFuncSetAttribute(cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cg2_kernel<<<2, 128, smem>>>(dA, dBk, dBm, dq, dp);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  std::vector<float> q(256 * 128), p(256 * 128);
  cudaMemcpy(q.data(), dq, q.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(p.data(), dp, p.size() * 4, cudaMemcpyDeviceToHost);
  double eq = 0, ep = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 128; ++n) {
      float rq = 0, rp = 0;
      for (int k = 0; k < KD; ++k) { rq += Af[m * KD + k] * Bkf[n * KD + k]; rp += Af[m * KD + k] * Bmf[k * 128 + n]; }
      eq = fmax(eq, fabs(rq - q[m * 128 + n]));
      ep = fmax(ep, fabs(rp - p[m * 128 + n]));
    }
  printf("cta_group::2 M=256 N=128: K-major B max err %.3g, MN-major B max err %.3g  (%s)\n", eq, ep, (eq == 0 && ep == 0) ? "OK" : "MISMATCH");
  if (eq != 0 || ep != 0) {
    printf("sample qk row0: "); for (int n = 0; n < 8; ++n) printf("%g ", q[n]); printf("| row0 n64..: "); for (int n = 64; n < 72; ++n) printf("%g ", q[n]);
    printf("| row128: "); for (int n = 0; n < 8; ++n) printf("%g ", q[128 * 128 + n]); printf("\n");
  }
  return 0;
}
