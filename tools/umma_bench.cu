// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16) as a function of N and of
// the operand layout, one CTA per SM, operands resident in shared memory (no loads in the loop).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu && ./umma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../anomaly_detection_on_video_b200/csrc/conv_umma.cuh"
#include "../anomaly_detection_on_video_b200/csrc/stem_umma.cuh"

using namespace vad;

template <int N, int ROW_BYTES>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, int per_commit, int a_step_bytes, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint64_t hi = umma_desc_kmajor<ROW_BYTES>(0);
    const uint32_t a0 = smem_u32(smem) >> 4;
    const uint32_t b0 = (smem_u32(smem) + 65536) >> 4;  // A region: 64 KB, B region: up to 32 KB + 16 KB of offsets
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
        for (int j = 0; j < per_commit; ++j) {
          const uint32_t off = (uint32_t)((j * a_step_bytes) & 16383) >> 4;
          umma_f16_c<true>(tmem, hi | (a0 + off), hi | (b0 + off), idesc);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// Pipelined variant: groups of `per_commit` MMAs, each group followed by a commit on a ring of 8 barriers and
// preceded by a wait on the barrier of the group issued `lag` groups earlier (what a real smem-ring consumer
// does); measures whether barrier traffic between groups stalls the MMA stream.
template <int N>
__global__ void __launch_bounds__(128, 1) bench_ring_kernel(int groups, int per_commit, int lag, int extra_polls, long long* out_cycles, int flags = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar[8];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&done_bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint64_t hi = umma_desc_kmajor<128>(0);
    const uint32_t a0 = smem_u32(smem) >> 4;
    const uint32_t b0 = (smem_u32(smem) + 65536) >> 4;
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (g >= lag && !(flags & 1)) mbar_wait(&bar[(g - lag) & 7], ((g - lag) >> 3) & 1);
      for (int e = 0; e < extra_polls; ++e) (void)mbar_try_wait(&done_bar, 1);  // already-complete phase: returns at once
      if (!(flags & 2)) tc_fence_after();
      if (elect_one_sync()) {
        for (int j = 0; j < per_commit; ++j) {
          const uint32_t off = (uint32_t)((j * 1024) & 16383) >> 4;
          umma_f16_c<true>(tmem, hi | (a0 + off), hi | (b0 + off), idesc);
        }
        if (!(flags & 4) || g >= groups - 8) umma_commit(&bar[g & 7]);
      }
      if (!(flags & 8)) __syncwarp();
    }
    if (flags & 4) { mbar_wait(&bar[(groups - 1) & 7], 0); } else
    for (int g = groups - lag; g < groups; ++g) mbar_wait(&bar[g & 7], (g >> 3) & 1);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}


// Lean ring: everything address-like precomputed, PER_COMMIT unrolled at compile time; what a tuned
// smem-ring consumer looks like.  NOSW: A through the stem's SWIZZLE_NONE descriptor (LBO 16 B, SBO 176 B).
template <int N, int PER_COMMIT, bool NOSW>
__global__ void __launch_bounds__(128, 1) bench_lean_kernel(int groups, int lag, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint64_t hiB = umma_desc_kmajor<128>(0);
    const uint64_t hiA = NOSW ? umma_desc_kmajor_noswizzle(0, 16u, 176u) : hiB;
    const uint32_t a0 = smem_u32(smem) >> 4;
    const uint32_t b0 = (smem_u32(smem) + 65536) >> 4;
    const uint32_t bar0 = smem_u32(&bar[0]);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (g >= lag) {
        const uint32_t addr = bar0 + (((uint32_t)(g - lag) & 7u) << 3), par = ((uint32_t)(g - lag) >> 3) & 1u;
        uint32_t ok, spins = 0;
        do {
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(addr), "r"(par) : "memory");
          if (++spins > (1u << 22)) __trap();
        } while (!ok);
      }
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int j = 0; j < PER_COMMIT; ++j) {
          const uint32_t off = NOSW ? (uint32_t)((j & 7) * 176) >> 4 : (uint32_t)((j * 1024) & 16383) >> 4;
          umma_f16_c<true>(tmem, hiA | (a0 + off), hiB | (b0 + (uint32_t)(j & 7) * 64), idesc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + (((uint32_t)g & 7u) << 3)) : "memory");
      }
      __syncwarp();
    }
    for (int g = groups - lag; g < groups; ++g) mbar_wait(&bar[g & 7], (g >> 3) & 1);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N, int PC, bool NOSW>
void run_lean(int lag) {
  long long* d;
  cudaMalloc(&d, 8);
  const int groups = 4000;
  cudaFuncSetAttribute(bench_lean_kernel<N, PC, NOSW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  bench_lean_kernel<N, PC, NOSW><<<148, 128, 164 * 1024>>>(16, lag, d);
  bench_lean_kernel<N, PC, NOSW><<<148, 128, 164 * 1024>>>(groups, lag, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("lean %s N=%3d per_commit=%2d lag=%d : %7.1f cycles/MMA, %7.1f cycles/group (%s)\n", NOSW ? "noswz" : "sw128", N, PC, lag,
         (double)c / ((double)groups * PC), (double)c / groups, cudaGetErrorString(e));
  cudaFree(d);
}


// Lean loop ablation.  FL bits: 1 no wait, 2 no fence, 4 no commit (except last 8), 8 no syncwarp,
// 16 whole loop inside one elected thread (no per-group elect / reconvergence)
template <int N, int PER_COMMIT, int FL>
__global__ void __launch_bounds__(128, 1) bench_abl_kernel(int groups, int lag, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint64_t hi = umma_desc_kmajor<128>(0);
    const uint32_t a0 = smem_u32(smem) >> 4;
    const uint32_t b0 = (smem_u32(smem) + 65536) >> 4;
    const uint32_t bar0 = smem_u32(&bar[0]);
    long long t0 = clock64();
    auto body = [&](int g) {
      if (!(FL & 1) && g >= lag) {
        const uint32_t addr = bar0 + (((uint32_t)(g - lag) & 7u) << 3), par = ((uint32_t)(g - lag) >> 3) & 1u;
        uint32_t ok, spins = 0;
        do {
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(addr), "r"(par) : "memory");
          if (++spins > (1u << 22)) __trap();
        } while (!ok);
      }
      if (!(FL & 2)) tc_fence_after();
    };
    auto issue = [&](int g) {
#pragma unroll
      for (int j = 0; j < PER_COMMIT; ++j) {
        const uint32_t off = (uint32_t)((j * 1024) & 16383) >> 4;
        umma_f16_c<true>(tmem, hi | (a0 + off), hi | (b0 + (uint32_t)(j & 7) * 64), idesc);
      }
      if (!(FL & 4) || g >= groups - 8)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + (((uint32_t)g & 7u) << 3)) : "memory");
    };
    if (FL & 16) {
      if (elect_one_sync()) {
        for (int g = 0; g < groups; ++g) { body(g); issue(g); }
      }
      __syncwarp();
    } else {
      for (int g = 0; g < groups; ++g) {
        body(g);
        if (elect_one_sync()) issue(g);
        if (!(FL & 8)) __syncwarp();
      }
    }
    if (FL & 4) mbar_wait(&bar[(groups - 1) & 7], 0);
    else for (int g = groups - lag; g < groups; ++g) mbar_wait(&bar[g & 7], (g >> 3) & 1);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N, int PC, int FL>
void run_abl(int lag = 4) {
  long long* d;
  cudaMalloc(&d, 8);
  const int groups = 4000;
  cudaFuncSetAttribute(bench_abl_kernel<N, PC, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  bench_abl_kernel<N, PC, FL & ~4><<<148, 128, 164 * 1024>>>(16, lag, d);
  bench_abl_kernel<N, PC, FL><<<148, 128, 164 * 1024>>>(groups, lag, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("abl FL=%2d N=%3d per_commit=%2d lag=%d : %7.1f cycles/MMA, %7.1f cycles/group (%s)\n", FL, N, PC, lag,
         (double)c / ((double)groups * PC), (double)c / groups, cudaGetErrorString(e));
  cudaFree(d);
}


// Two-warp pipeline like the real kernels: warp 0 "producer" waits empty[s] (signalled by tcgen05.commit) and
// arrives full[s] (stands in for the TMA completion); warp 1 waits full[s], issues PER_COMMIT MMAs, commits empty[s].
template <int N, int PER_COMMIT, int STAGES>
__global__ void __launch_bounds__(128, 1) bench_pipe_kernel(int groups, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t full[STAGES], empty[STAGES];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
  auto wait = [](uint32_t addr, uint32_t par) {
    uint32_t ok, spins = 0;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(addr), "r"(par) : "memory");
      if (++spins > (1u << 22)) __trap();
    } while (!ok);
  };
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    for (int g = 0; g < groups; ++g) {
      wait(empty0 + s * 8, ph ^ 1);
      if (elect_one_sync()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + s * 8) : "memory");
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint64_t hi = umma_desc_kmajor<128>(0);
    const uint32_t a0 = smem_u32(smem) >> 4;
    const uint32_t b0 = (smem_u32(smem) + 65536) >> 4;
    long long t0 = clock64();
    uint32_t s = 0, ph = 0;
    for (int g = 0; g < groups; ++g) {
      wait(full0 + s * 8, ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int j = 0; j < PER_COMMIT; ++j) {
          const uint32_t off = (uint32_t)((j * 1024) & 16383) >> 4;
          umma_f16_c<true>(tmem, hi | (a0 + off), hi | (b0 + (uint32_t)(j & 7) * 64), idesc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + s * 8) : "memory");
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    // drain: the last commit
    { uint32_t ls = (uint32_t)((groups - 1) % STAGES), lp = (uint32_t)(((groups - 1) / STAGES) & 1); wait(empty0 + ls * 8, lp); }
    long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N, int PC, int STAGES>
void run_pipe() {
  long long* d;
  cudaMalloc(&d, 8);
  const int groups = 4000;
  cudaFuncSetAttribute(bench_pipe_kernel<N, PC, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  bench_pipe_kernel<N, PC, STAGES><<<148, 128, 164 * 1024>>>(16, d);
  bench_pipe_kernel<N, PC, STAGES><<<148, 128, 164 * 1024>>>(groups, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("pipe N=%3d per_commit=%2d stages=%d : %7.1f cycles/MMA, %7.1f cycles/group (%s)\n", N, PC, STAGES,
         (double)c / ((double)groups * PC), (double)c / groups, cudaGetErrorString(e));
  cudaFree(d);
}


// Software-pipelined issuer: one elected thread owns the whole loop; the wait for stage s+1 is placed between the
// MMAs of stage s (after WAIT_AFTER of them), so its latency overlaps queued tensor work.
template <int N, int PER_COMMIT, int STAGES, int WAIT_AFTER>
__global__ void __launch_bounds__(128, 1) bench_swp_kernel(int groups, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t full[STAGES], empty[STAGES];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
  auto wait = [](uint32_t addr, uint32_t par) {
    uint32_t ok, spins = 0;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(addr), "r"(par) : "memory");
      if (++spins > (1u << 22)) __trap();
    } while (!ok);
  };
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    for (int g = 0; g < groups; ++g) {
      wait(empty0 + s * 8, ph ^ 1);
      if (elect_one_sync()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + s * 8) : "memory");
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(N);
      const uint64_t hi = umma_desc_kmajor<128>(0);
      const uint32_t a0 = smem_u32(smem) >> 4;
      const uint32_t b0 = (smem_u32(smem) + 65536) >> 4;
      long long t0 = clock64();
      uint32_t s = 0, ph = 0;
      wait(full0, 0);
      tc_fence_after();
      for (int g = 0; g < groups; ++g) {
        uint32_t ns = s + 1, nph = ph;
        if (ns == STAGES) { ns = 0; nph ^= 1; }
#pragma unroll
        for (int j = 0; j < PER_COMMIT; ++j) {
          const uint32_t off = (uint32_t)((j * 1024) & 16383) >> 4;
          umma_f16_c<true>(tmem, hi | (a0 + off), hi | (b0 + (uint32_t)(j & 7) * 64), idesc);
          if (j == WAIT_AFTER - 1 && g + 1 < groups) { wait(full0 + ns * 8, nph); tc_fence_after(); }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + s * 8) : "memory");
        s = ns; ph = nph;
      }
      { uint32_t ls = (uint32_t)((groups - 1) % STAGES), lp = (uint32_t)(((groups - 1) / STAGES) & 1); wait(empty0 + ls * 8, lp); }
      long long t1 = clock64();
      if (blockIdx.x == 0) *out_cycles = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N, int PC, int STAGES, int WA>
void run_swp() {
  long long* d;
  cudaMalloc(&d, 8);
  const int groups = 4000;
  cudaFuncSetAttribute(bench_swp_kernel<N, PC, STAGES, WA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  bench_swp_kernel<N, PC, STAGES, WA><<<148, 128, 164 * 1024>>>(16, d);
  bench_swp_kernel<N, PC, STAGES, WA><<<148, 128, 164 * 1024>>>(groups, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("swp N=%3d per_commit=%2d stages=%d wait_after=%d : %7.1f cycles/MMA, %7.1f cycles/group (%s)\n", N, PC, STAGES, WA,
         (double)c / ((double)groups * PC), (double)c / groups, cudaGetErrorString(e));
  cudaFree(d);
}


// Stem-like operands: A through the SWIZZLE_NONE descriptor (LBO 16 B, SBO 176 B), B with BROW-byte swizzled rows,
// N in {64, 128, 192}; PER_COMMIT MMAs per group, no waits in the loop (pure issue/execute rate).
template <int N, int PER_COMMIT, int BROW, bool ANOSW>
__global__ void __launch_bounds__(128, 1) bench_stemlike_kernel(int groups, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(N);
      const uint64_t hiB = umma_desc_kmajor<BROW>(0);
      const uint64_t hiA = ANOSW ? umma_desc_kmajor_noswizzle(0, 16u, 176u) : umma_desc_kmajor<64>(0);
      const uint32_t a0 = smem_u32(smem) >> 4;
      const uint32_t b0 = (smem_u32(smem) + 65536) >> 4;
      long long t0 = clock64();
      for (int g = 0; g < groups; ++g) {
#pragma unroll
        for (int j = 0; j < PER_COMMIT; ++j) {
          const uint32_t aoff = ANOSW ? (uint32_t)(((j >> 1) & 3) * 176 + ((j >> 1) & 4 ? 3456 : 0)) >> 4 : (uint32_t)((j >> 1) * 1024) >> 4;
          const uint32_t boff = (uint32_t)((j >> 1) * (N * BROW)) >> 4;
          umma_f16_c<true>(tmem, hiA | (a0 + aoff + 2 * (j & 1)), hiB | (b0 + (boff & 4095) + 2 * (j & 1)), idesc);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) *out_cycles = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int PC, int BROW, bool ANOSW>
void run_stemlike() {
  long long* d;
  cudaMalloc(&d, 8);
  const int groups = 2000;
  cudaFuncSetAttribute(bench_stemlike_kernel<N, PC, BROW, ANOSW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  bench_stemlike_kernel<N, PC, BROW, ANOSW><<<148, 128, 164 * 1024>>>(16, d);
  bench_stemlike_kernel<N, PC, BROW, ANOSW><<<148, 128, 164 * 1024>>>(groups, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("stemlike N=%3d B rows %3d B, A %s : %7.1f cycles/MMA (%s)\n", N, BROW, ANOSW ? "noswizzle(16,176)" : "SW64", (double)c / ((double)groups * PC),
         cudaGetErrorString(e));
  cudaFree(d);
}

template <int N>
void run_ring(int per_commit, int lag, int extra_polls, int flags = 0) {
  long long* d;
  cudaMalloc(&d, 8);
  const int groups = 4000;
  cudaFuncSetAttribute(bench_ring_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  bench_ring_kernel<N><<<148, 128, 164 * 1024>>>(16, per_commit, lag, extra_polls, d, flags & ~4);
  bench_ring_kernel<N><<<148, 128, 164 * 1024>>>(groups, per_commit, lag, extra_polls, d, flags);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("ring flags=%2d N=%3d per_commit=%2d lag=%d extra_polls=%d : %7.1f cycles/MMA, %7.1f cycles/group (%s)\n", flags, N, per_commit, lag,
         extra_polls, (double)c / ((double)groups * per_commit), (double)c / groups, cudaGetErrorString(e));
  cudaFree(d);
}

template <int N, int ROW_BYTES>
void run(const char* name, int per_commit, int a_step) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 2000;
  cudaFuncSetAttribute(bench_kernel<N, ROW_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  bench_kernel<N, ROW_BYTES><<<148, 128, 164 * 1024>>>(10, per_commit, a_step, d);
  bench_kernel<N, ROW_BYTES><<<148, 128, 164 * 1024>>>(iters, per_commit, a_step, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d row=%3dB per_commit=%3d a_step=%5d : %7.1f cycles/MMA  (%s)\n", name, N, ROW_BYTES, per_commit, a_step,
         (double)c / ((double)iters * per_commit), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  run_stemlike<64,14,64,true>(); run_stemlike<128,14,64,true>(); run_stemlike<192,14,64,true>(); run_stemlike<64,14,128,true>(); run_stemlike<128,14,128,true>(); run_stemlike<192,14,128,true>(); run_stemlike<128,14,64,false>(); run_stemlike<192,14,64,false>(); run_stemlike<256,14,64,true>();
  return 0;
  run_swp<64,4,4,1>(); run_swp<64,4,4,2>(); run_swp<64,4,4,3>(); run_swp<64,4,4,4>(); run_swp<64,8,4,2>(); run_swp<64,8,4,4>(); run_swp<64,8,4,6>(); run_swp<128,4,4,1>(); run_swp<128,4,4,2>(); run_swp<128,4,4,3>(); run_swp<128,8,4,4>(); run_swp<256,4,4,2>(); run_swp<64,14,4,7>(); run_swp<64,2,4,1>();
  run_pipe<64,4,4>(); run_pipe<64,4,8>(); run_pipe<64,2,8>(); run_pipe<64,8,4>(); run_pipe<64,14,4>(); run_pipe<128,4,4>(); run_pipe<128,4,8>(); run_pipe<128,8,4>(); run_pipe<256,4,4>(); run_pipe<256,2,4>(); run_pipe<64,4,2>(); run_pipe<64,4,3>();
  return 0;
  run_abl<64,4,0>(); run_abl<64,4,1>(); run_abl<64,4,2>(); run_abl<64,4,8>(); run_abl<64,4,5>(); run_abl<64,4,7>(); run_abl<64,4,15>(); run_abl<64,4,16>(); run_abl<64,4,17>(); run_abl<64,4,21>(); run_abl<64,4,23>();
  run_abl<64,8,16>(); run_abl<64,14,16>(); run_abl<128,4,16>(); run_abl<128,8,16>(); run_abl<256,2,16>(); run_abl<64,1,23>(); run_abl<64,2,23>(); run_abl<128,1,23>(); run_abl<256,1,23>();
  run_lean<64, 2, false>(4); run_lean<64, 4, false>(4); run_lean<64, 8, false>(4); run_lean<64, 14, false>(4); run_lean<64, 64, false>(4);
  run_lean<64, 4, true>(4); run_lean<64, 14, true>(4); run_lean<64, 64, true>(4);
  run_lean<128, 4, false>(4); run_lean<128, 8, false>(4); run_lean<256, 4, false>(4); run_lean<256, 2, false>(4);
  run_lean<64, 4, false>(1); run_lean<64, 4, false>(2);
  return 0;
  for (int pc : {4, 14, 64}) {
    run<64, 128>("SW128 same operands", pc, 0);
    run<64, 128>("SW128 A advances 32B", pc, 32);
    run<64, 128>("SW128 A advances 1KB", pc, 1024);
    run<128, 128>("SW128 same operands", pc, 0);
    run<128, 128>("SW128 A advances 1KB", pc, 1024);
    run<256, 128>("SW128 same operands", pc, 0);
    run<256, 128>("SW128 A advances 1KB", pc, 1024);
    run<64, 64>("SW64 A advances 1KB", pc, 1024);
  }
  for (int pc : {2, 4, 8, 14}) {
    for (int lag : {1, 2, 4, 6}) run_ring<64>(pc, lag, 0);
    run_ring<64>(pc, 4, 2);
  }
  for (int pc : {4, 8}) { run_ring<128>(pc, 4, 0); run_ring<256>(pc, 4, 0); }
  printf("flags: 1 = no wait, 2 = no tcgen05.fence, 4 = no commit, 8 = no syncwarp\n");
  for (int f : {1, 2, 8, 3, 5, 7, 15}) run_ring<64>(4, 4, 0, f);
  return 0;
}
