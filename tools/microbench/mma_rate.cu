// Issue rate of back-to-back tcgen05.mma (bf16, K-major SW128 operands already in shared memory, no TMA, no
// epilogue): cycles per MMA instruction for the shapes the scan kernel uses.  Diagnostics only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../knowledge_enhanced_multimodal_retrieval_b200/csrc mma_rate.cu
#include <cstdio>
#include "scan_mma.cuh"
using namespace kemr;

// mode: N per MMA; alt = alternate between two accumulators; PAIR = cta_group::2 (M = 256)
template <bool PAIR>
__global__ void __launch_bounds__(320, 1) mma_rate_kernel(int N, int alt, int iters, long long* out, int commit_every, int ldx, int gap) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2[8];
  __shared__ volatile int done;
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) ptx::mbar_init(&bar2[i], 1); done = 0; ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { if (PAIR) ptx::tmem_alloc_pair(&tmem_ptr, 512); else ptx::tmem_alloc(&tmem_ptr, 512); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, N);
    const uint32_t sa = ptx::smem_u32(smem);
    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + 16 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + ((alt && (it & 1)) ? 256u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (PAIR) ptx::mma_bf16_pair(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
        else ptx::mma_bf16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
      }
      if (commit_every && (it % commit_every) == commit_every - 1) {
        if (PAIR) ptx::mma_commit_pair(&bar2[it & 7], (uint16_t)3); else ptx::mma_commit(&bar2[it & 7]);
      }
    }
    if (PAIR) ptx::mma_commit_pair(&bar, (uint16_t)1); else ptx::mma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    done = 1;
  } else if (threadIdx.x >= 64 && ldx) {
    // epilogue-like readers of the OTHER accumulator buffer (columns 256..511), lane quadrant warp % 4
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    uint32_t sink = 0;
    while (!done && !(PAIR && rank)) {
      for (int c = 0; c < 128 && !done; c += ldx) {
        if (ldx == 16) { uint32_t r[16]; tmem_ld16(lane_addr + ((warp - 2) >> 2) * 128 + c, r); ptx::tmem_ld_wait(); sink += r[3]; }
        else { uint32_t r[32]; ptx::tmem_ld32(lane_addr + ((warp - 2) >> 2) * 128 + c, r); ptx::tmem_ld_wait(); sink += r[7]; }
        const long long t = clock64();
        while (clock64() - t < gap) {}
      }
    }
    if (sink == 0x12345) out[0] = sink;
  }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); if (PAIR) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512); }
}

template <bool PAIR>
void run(int N, int alt, int grid, int commit_every = 0, int ldx = 0, int gap = 0) {
  const int iters = 2000, smem = 64 * 1024;
  long long* out; cudaMalloc(&out, grid * 8); cudaMemset(out, 0, grid * 8);
  cudaFuncSetAttribute(mma_rate_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchKernelEx(&cfg, mma_rate_kernel<PAIR>, N, alt, iters, out, commit_every, ldx, gap);
  cudaEventRecord(a);
  cudaLaunchKernelEx(&cfg, mma_rate_kernel<PAIR>, N, alt, iters, out, commit_every, ldx, gap);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  long long h[512]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; int n = 0;
  for (int i = 0; i < grid; i += PAIR ? 2 : 1) { avg += (double)h[i]; ++n; }
  avg /= n;
  const double flop = 2.0 * (PAIR ? 256 : 128) * N * 16 * 4.0 * iters * n;
  printf("%s M=%d N=%3d alt=%d commit/%d ldx=%d gap=%d grid=%3d: %.1f cycles/MMA (ideal %.0f), %.3f ms, %.0f TFLOP/s  %s\n", PAIR ? "pair" : "cta ",
         PAIR ? 256 : 128, N, alt, commit_every, ldx, gap, grid, avg / (4.0 * iters), N / 2.0, ms, flop / (ms * 1e-3) / 1e12, cudaGetErrorString(e));
  cudaFree(out);
}

// The scan kernel's stage ring without any data movement: the MMA thread waits full[s], issues 4 MMAs, commits to
// empty[s]; a producer thread waits empty[s], waits `lat` more cycles (stand-in for the TMA latency), arrives full[s].
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) mma_ring_kernel(int S, int lat, int iters, long long* out, int slot_stride, int same_a) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, full[16], empty[16];
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    for (int i = 0; i < 16; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) { if (PAIR) ptx::tmem_alloc_pair(&tmem_ptr, 512); else ptx::tmem_alloc(&tmem_ptr, 512); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 32 && rank == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      ptx::mbar_wait(&empty[s], ph ^ 1);
      if (lat) { const long long t = clock64(); while (clock64() - t < lat) {} }
      ptx::mbar_arrive(&full[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, 256);
    const uint32_t sa = ptx::smem_u32(smem);
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      ptx::mbar_wait(&full[s], ph);
      ptx::tc_fence_after();
      const uint32_t so = sa + (uint32_t)(s * slot_stride);
      const uint64_t adesc = umma_desc_sw128(same_a ? sa : so), bdesc = umma_desc_sw128(so + 16 * 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (PAIR) ptx::mma_bf16_pair(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
        else ptx::mma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
      }
      if (PAIR) ptx::mma_commit_pair(&empty[s], (uint16_t)3); else ptx::mma_commit(&empty[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (PAIR) ptx::mma_commit_pair(&bar, (uint16_t)1); else ptx::mma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); if (PAIR) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512); }
}

template <bool PAIR>
void run_ring(int S, int lat, int slot_stride = 0, int same_a = 0) {
  const int iters = 2000, smem = 202 * 1024, grid = 148;
  long long* out; cudaMalloc(&out, grid * 8); cudaMemset(out, 0, grid * 8);
  cudaFuncSetAttribute(mma_ring_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, mma_ring_kernel<PAIR>, S, lat, iters, out, slot_stride, same_a);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[512]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; int n = 0;
  for (int i = 0; i < grid; i += PAIR ? 2 : 1) { avg += (double)h[i]; ++n; }
  printf("ring %s stages=%2d lat=%4d stride=%d sameA=%d: %.1f cycles/MMA (ideal 128)  %s\n", PAIR ? "pair" : "cta ", S, lat, slot_stride, same_a, avg / n / (4.0 * iters), cudaGetErrorString(e));
  cudaFree(out);
}

int main_rate();
int main() {
  run_ring<true>(6, 0, 0, 0);
  run_ring<true>(6, 0, 32 * 1024, 0);
  run_ring<true>(6, 0, 32 * 1024, 1);
  run_ring<true>(2, 0, 32 * 1024, 0);
  run_ring<false>(4, 0, 48 * 1024, 0);
  run_ring<false>(4, 0, 48 * 1024, 1);
  return main_rate();
}
int main_rate() {
  for (int grid : {1, 148}) {
    for (int N : {64, 128, 256}) run<false>(N, 0, grid);
    run<false>(256, 1, grid);
  }
  for (int grid : {2, 148}) {
    for (int N : {128, 256}) run<true>(N, 0, grid);
    run<true>(256, 1, grid);
  }
  // what the scan kernel adds around the MMAs: a commit per 4 MMAs, and epilogue warps reading the other buffer
  run<true>(256, 0, 148, 1, 0, 0);
  for (int ldx : {16, 32}) for (int gap : {0, 500, 2000}) run<true>(256, 0, 148, 0, ldx, gap);
  run<true>(256, 0, 148, 1, 16, 500);
  run<false>(256, 0, 148, 1, 16, 500);
  return 0;
}
