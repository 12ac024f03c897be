// Microbenchmark: sustained issue rate of tcgen05.mma.kind::f16 (cta_group::1, M = 128, K = 16) as a function of N,
// with the A operand in tensor memory (TS) or in shared memory (SS).  One CTA per SM, every CTA issues `n_mma`
// back-to-back MMAs into one accumulator and waits for the commit; prints cycles per MMA of CTA 0 and the slowest CTA.
//   mma_rate_probe [n_mma]
// Developer tool (not part of the product library); operands are whatever shared / tensor memory holds.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../jaeger_b200/csrc/conv_ws.cuh"

using namespace jg::tc;

// variant bits: 1 = rotate over 3 accumulators every 40 MMAs (first MMA of a group overwrites), 2 = commit after every group of 40,
// 4 = two other warps read the accumulators with tcgen05.ld in a loop meanwhile, 8 = another warp streams bulk copies into shared memory
template <bool kTS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int n_mma, int b_rows_step, int variant, const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_init(smem_u32(&bar3), 1); s_done = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512u);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t b_lo = desc_lo_sw128(smem_u32(smem));
    const uint32_t a_lo = desc_lo_sw128(smem_u32(smem) + 32768u);
    const long long t0 = clock64();
    const int n_groups = n_mma / 40;
    for (int gi = 0; gi < n_groups; ++gi) {
      const uint32_t d = tmem + 320u + ((variant & 1) ? static_cast<uint32_t>((gi % 3) * 64) : 0u);
      if (leader) {
#pragma unroll
        for (int j = 0; j < 40; ++j) {
          const uint32_t k16 = static_cast<uint32_t>(j & 3);
          const uint32_t b = b_lo + static_cast<uint32_t>(((j >> 2) % 5) * b_rows_step * 8) + k16 * 2u;
          const uint32_t acc = (variant & 1) ? (j != 0) : ((gi | j) != 0);
          if (kTS) jg::ws::umma_f16_ts(d, tmem + static_cast<uint32_t>(j * 8), desc_pack(b, kDescHiSw128), idesc, acc);
          else umma_bf16(d, desc_pack(a_lo + k16 * 2u, kDescHiSw128), desc_pack(b, kDescHiSw128), idesc, acc);
        }
        if (variant & 2) umma_commit(smem_u32(&bar2));
      }
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; s_done = 1; }
  } else if ((warp == 2 || warp == 3) && (variant & 4)) {
    uint32_t v[16];
    uint32_t sink = 0;
    while (!s_done) {
      jg::ws::tmem_ld16x256_x4(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 320u, v);
      tmem_ld_wait();
      sink += v[0] + v[15];
    }
    if (sink == 0x12345u) out[0] = 0;
  } else if (warp == 0 && (variant & 8)) {
    uint32_t ph = 0;
    while (!s_done) {
      if (lane == 0) {
        mbar_expect_tx(smem_u32(&bar3), 18432u);
        bulk_g2s(smem_u32(smem) + 49152u, gsrc + (static_cast<size_t>(blockIdx.x) * 64 + (ph & 63)) * 18432u, 18432u, smem_u32(&bar3));
      }
      mbar_wait(smem_u32(&bar3), ph & 1u);
      ++ph;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512u); }
}

int main(int argc, char** argv) {
  const int n_mma = argc > 1 ? atoi(argv[1]) : 4000;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * 8);
  std::vector<long long> h(sms);
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  uint8_t* gsrc;
  cudaMalloc(&gsrc, static_cast<size_t>(sms) * 64 * 18432);
  cudaMemset(gsrc, 0, static_cast<size_t>(sms) * 64 * 18432);
  for (int variant : {0, 1, 2, 3, 4, 8, 15})
  for (int ts = 1; ts >= 0; --ts)
    for (int n : {64, 128}) {
      if (variant && (!ts || n != 64)) continue;
      for (int rep = 0; rep < 2; ++rep) {
        if (ts) rate_kernel<true><<<sms, 128, smem>>>(n, n_mma, 3, variant, gsrc, d);
        else rate_kernel<false><<<sms, 128, smem>>>(n, n_mma, 3, variant, gsrc, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h.data(), d, sms * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (long long v : h) mx = v > mx ? v : mx;
      printf("variant %2d %s M=128 N=%3d K=16: %.1f cycles/MMA (CTA 0), %.1f (slowest CTA); ideal %d\n", variant, ts ? "TS" : "SS", n,
             double(h[0]) / n_mma, double(mx) / n_mma, n / 2);
    }
  return 0;
}
