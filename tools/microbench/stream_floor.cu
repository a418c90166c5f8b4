// What does a 132 MB read cost on this GPU?  The HBM roofline denominator (MEASURED_PEAKS.json) is a 2 GiB copy, where
// launch, ramp-up and tail vanish; BASELINE config 3 streams only 132 MB (43 000 rows x 768-d bf16 x 2 galleries), so
// the practical floor of ANY kernel that reads those bytes once is what this program measures:
//   ldg   : 2 CTAs x 256 threads per SM, contiguous range per CTA, 8 independent 16-byte loads in flight per thread
//   bulk  : 1 CTA per SM, a producer warp streams the CTA's range through a ring of shared-memory stages with
//           cp.async.bulk (UBLKCP) + mbarriers, 8 consumer warps only wait and release (no arithmetic)
// for 132 MB and 2 GiB, L2 flushed (512 MiB memset) before every launch, CUDA events, median of 20.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256, 2) read_ldg(const uint4* __restrict__ p, size_t n16, uint32_t* out) {
  const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t b = per * blockIdx.x, e = b + per < n16 ? b + per : n16;
  uint32_t acc = 0;
  size_t i = b + threadIdx.x;
  for (; i + 7 * 256 < e; i += 8 * 256) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * 256));
#pragma unroll
    for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < e; i += 256) { const uint4 v = p[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) out[blockIdx.x] = acc;       // keeps the loads alive
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int STAGE_BYTES, int STAGES>
__global__ void __launch_bounds__(288, 1) read_bulk(const unsigned char* __restrict__ p, size_t bytes, uint32_t* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per = ((bytes + gridDim.x - 1) / gridDim.x + 127) / 128 * 128;
  const size_t b = per * blockIdx.x, e = b + per < bytes ? b + per : bytes;
  const int niter = b < e ? (int)((e - b + STAGE_BYTES - 1) / STAGE_BYTES) : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int stage = 0; uint32_t phase = 0;
  if (warp == 8) {
    for (int it = 0; it < niter; ++it) {
      while (!try_wait(&empty[stage], phase ^ 1)) {}
      if (lane == 0) {
        const size_t off = b + (size_t)it * STAGE_BYTES;
        const uint32_t n = (uint32_t)(e - off < (size_t)STAGE_BYTES ? e - off : (size_t)STAGE_BYTES);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stage])), "r"(n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + (size_t)stage * STAGE_BYTES)), "l"(p + off), "r"(n), "r"(smem_u32(&full[stage])) : "memory");
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    uint32_t acc = 0;
    for (int it = 0; it < niter; ++it) {
      while (!try_wait(&full[stage], phase)) {}
      acc ^= reinterpret_cast<const uint32_t*>(smem + (size_t)stage * STAGE_BYTES)[threadIdx.x];
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (acc == 0x12345678u) out[blockIdx.x] = acc;
  }
}

template <class F>
static float median_us(F launch, void* flush, size_t flush_bytes, int reps = 20) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  std::vector<float> t;
  for (int i = 0; i < reps + 3; ++i) {
    CK(cudaMemsetAsync(flush, i, flush_bytes));
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (i >= 3) t.push_back(ms * 1e3f);
  }
  std::sort(t.begin(), t.end());
  return t[t.size() / 2];
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t flush_bytes = (size_t)512 << 20;
  void* flush; CK(cudaMalloc(&flush, flush_bytes));
  uint32_t* out; CK(cudaMalloc(&out, 4096 * 4));
  const size_t sizes[2] = {(size_t)43000 * 768 * 2 * 2, (size_t)2 << 30};
  for (size_t bytes : sizes) {
    unsigned char* buf; CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    const float t1 = median_us([&] { read_ldg<<<2 * sms, 256>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, out); }, flush, flush_bytes);
    printf("%8.1f MB  ldg  (2 x 256 threads / SM, 8 x 16 B in flight per thread): %8.2f us  %7.1f GB/s\n", bytes / 1e6, t1, bytes / t1 / 1e3);
    {
      constexpr int SB = 48 * 1024, ST = 4;
      const size_t smem = (size_t)SB * ST + 256;
      CK(cudaFuncSetAttribute(read_bulk<SB, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const float t2 = median_us([&] { read_bulk<SB, ST><<<sms, 288, smem>>>(buf, bytes, out); }, flush, flush_bytes);
      printf("%8.1f MB  bulk (1 CTA / SM, 4 stages x 48 KB):                          %8.2f us  %7.1f GB/s\n", bytes / 1e6, t2, bytes / t2 / 1e3);
    }
    {
      constexpr int SB = 16 * 1024, ST = 12;
      const size_t smem = (size_t)SB * ST + 256;
      CK(cudaFuncSetAttribute(read_bulk<SB, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const float t2 = median_us([&] { read_bulk<SB, ST><<<sms, 288, smem>>>(buf, bytes, out); }, flush, flush_bytes);
      printf("%8.1f MB  bulk (1 CTA / SM, 12 stages x 16 KB):                         %8.2f us  %7.1f GB/s\n", bytes / 1e6, t2, bytes / t2 / 1e3);
    }
    {
      const float t3 = median_us([&] { read_ldg<<<1, 32>>>(reinterpret_cast<const uint4*>(buf), 32, out); }, flush, flush_bytes);
      printf("            empty-ish launch (1 warp, 512 B):                                  %8.2f us\n", t3);
    }
    CK(cudaFree(buf));
  }
  return 0;
}
