// ubench_tmem.cu — micro-benchmarks that size the attention softmax loop on sm_100a:
//   (1) tcgen05.ld throughput / latency as a function of the number of warps,
//   (2) tcgen05.st throughput,
//   (3) MUFU.EX2 throughput,
//   (4) the fused "ld S -> exp2 -> pack bf16 -> st P" loop of the forward kernel.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/ubench_tmem tools/ubench_tmem.cu
// One CTA on one SM; cycles are clock64() deltas of the slowest warp.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vision_transformers_torch_xla_b200/csrc/vitk_common.cuh"

using namespace vitk;

// mode 0: ld.x32 back to back, one wait per 4 loads     (throughput)
// mode 1: ld.x32 + wait every load                       (latency-bound)
// mode 2: st.x8 back to back                             (store throughput)
// mode 3: ex2 only, 32 independent per iteration         (MUFU throughput)
// mode 4: fused: ld.x16 -> 16x(ffma, ex2), sum -> pack -> st.x8, software-pipelined by one chunk
// mode 5: like 4 but without the TMEM store (P kept in registers -> sink)
// mode 6: max pass: ld.x32 x2 -> 64 fmax
__global__ void __launch_bounds__(1024, 1) ubench(int mode, int iters, long long* cycles, float* sink, int workers, int comode = -1) {
  __shared__ uint32_t slot;
  __shared__ uint64_t spin_bar;
  if (threadIdx.x == 0) {
    mbar_init(&spin_bar, workers);
    fence_mbar_init();
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // initialise all 512 columns of this warp's lanes with small finite values
  if (warp < 4) {
    for (int c = 0; c < 512; c += 8) {
      uint32_t z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = __float_as_uint(0.001f * (float)((c + i + threadIdx.x) & 63));
      tmem_st_32x8(base + c, z);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // each warp works on its own 128-column window when there are several warps per lane quarter
  const uint32_t win = base + ((warp >> 2) & 3) * 128;
  float acc = 0.f;
  if (warp >= workers) {
    if (comode < 0) {
      // spinner: waits (like an idle softmax group / MMA warp) until every worker warp has finished
      mbar_wait(&spin_bar, 0);
      mode = -1;
    } else {
      iters *= 2;
      mode = comode;  // co-runner: a different phase of the other softmax group running at the same time
    }
  }
  const long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld_32x32(win + 0, a);
      tmem_ld_32x32(win + 32, b);
      tmem_ld_32x32(win + 64, c);
      tmem_ld_32x32(win + 96, d);
      tmem_ld_wait();
      acc += __uint_as_float(a[0] ^ a[31] ^ b[0] ^ b[31] ^ c[0] ^ c[31] ^ d[0] ^ d[31]);
    }
  } else if (mode == 1) {
    for (int it = 0; it < iters; ++it) {
      uint32_t a[32];
      tmem_ld_32x32(win + (it & 3) * 32, a);
      tmem_ld_wait();
      acc += __uint_as_float(a[0] ^ a[31]);
    }
  } else if (mode == 2) {
    uint32_t z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) tmem_st_32x8(win + u * 8, z);
    }
    tmem_st_wait();
  } else if (mode == 3) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = -0.001f * (float)(i + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = ex2_approx(x[i] - 1.0f);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += x[i];
  } else if (mode == 4 || mode == 5) {
    const float c2 = 0.18f, mc = 0.05f;
    float s0 = 0.f, s1 = 0.f;
    for (int it = 0; it < iters; ++it) {
      uint32_t cur[16], nxt[16];
      tmem_ld_32x16(win, cur);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // 8 chunks of 16 columns = one 128-column row segment
        if (u + 1 < 8) tmem_ld_32x16(win + (u + 1) * 16, nxt);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(cur[2 * i]), c2, -mc));
          const float e1 = ex2_approx(fmaf(__uint_as_float(cur[2 * i + 1]), c2, -mc));
          s0 += e0;
          s1 += e1;
          pk[i] = pack_bf16x2(e0, e1);
        }
        if (mode == 4) {
          tmem_st_32x8(win + u * 8, pk);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += __uint_as_float(pk[i]);
        }
        if (u + 1 < 8) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }
      }
    }
    if (mode == 4) tmem_st_wait();
    acc += s0 + s1;
  } else if (mode == 7) {
    // the rolled loop of attn_fwd3 pass 2 (7 units of 16 columns, register rotation by copy)
    const float c2 = 0.18f, mc = 0.05f;
    float s0 = 0.f, s1 = 0.f;
    const int my_k = 7 + (iters >> 20);
    for (int it = 0; it < iters; ++it) {
      uint32_t cur[16], nxt[16];
      tmem_ld_32x16(win, cur);
      tmem_ld_wait();
#pragma unroll 1
      for (int u = 0; u < my_k; ++u) {
        const bool more = u + 1 < my_k;
        if (more) tmem_ld_32x16(win + (u + 1) * 16, nxt);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(cur[2 * i]), c2, -mc));
          const float e1 = ex2_approx(fmaf(__uint_as_float(cur[2 * i + 1]), c2, -mc));
          s0 += e0;
          s1 += e1;
          pk[i] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x8(win + u * 8, pk);
        if (more) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }
      }
      tmem_st_wait();
    }
    acc += s0 + s1;
  } else if (mode == 6) {
    float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, m3 = -1e30f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t a[32], b[32];
        tmem_ld_32x32(win + h * 64, a);
        tmem_ld_32x32(win + h * 64 + 32, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = fmaxf(m0, __uint_as_float(a[i]));
          m1 = fmaxf(m1, __uint_as_float(a[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(b[i]));
          m3 = fmaxf(m3, __uint_as_float(b[i + 1]));
        }
      }
    }
    acc += fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }
  const long long t1 = clock64();
  if (warp < workers) {
    if ((threadIdx.x & 31) == 0) {
      cycles[warp] = t1 - t0;
      mbar_arrive(&spin_bar);
    }
  }
  if (acc == 12345.678f) sink[threadIdx.x] = acc;

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(slot, 512);
  }
}

int main() {
  long long* d_cycles;
  float* d_sink;
  cudaMalloc(&d_cycles, 32 * sizeof(long long));
  cudaMalloc(&d_sink, 1024 * sizeof(float));
  const int iters = 256;
  const char* names[] = {"ld.x32 x4 per wait (128 cols = 16 KB/warp/iter)",
                         "ld.x32 + wait (4 KB/warp/iter)",
                         "st.x8 x16 (128 cols = 16 KB/warp/iter)",
                         "ex2 x32 per thread per iter",
                         "fused ld->exp2->pack->st, 128 cols/thread/iter",
                         "fused ld->exp2->pack (no st), 128 cols/thread/iter",
                         "max pass, 128 cols/thread/iter",
                         "fused rolled loop (attn_fwd3 pass 2), 112 cols/thread/iter"};
  for (int mode = 0; mode < 8; ++mode) {
    printf("== mode %d: %s\n", mode, names[mode]);
    for (int warps : {1, 4, 8, 16, 32}) {
      ubench<<<1, warps * 32, 0>>>(mode, iters, d_cycles, d_sink, warps);  // warm-up
      ubench<<<1, warps * 32, 0>>>(mode, iters, d_cycles, d_sink, warps);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
      }
      std::vector<long long> h(32);
      cudaMemcpy(h.data(), d_cycles, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
      const double per_iter = (double)mx / iters;
      double elems = 0;  // 32-bit elements (or exps) processed by the whole CTA per iteration
      if (mode == 0 || mode == 2 || mode == 4 || mode == 5 || mode == 6) elems = warps * 32.0 * 128;
      if (mode == 1 || mode == 3) elems = warps * 32.0 * 32;
      if (mode == 7) elems = warps * 32.0 * 112;
      printf("   warps=%2d  %9.1f cyc/iter   %7.2f elem/cyc/SM  (%7.1f B/cyc)\n", warps, per_iter, elems / per_iter,
             4 * elems / per_iter);
    }
  }
  // spinners: 8 worker warps run the fused loop while S extra warps spin on an mbarrier (idle group / control warps)
  for (int mode : {4, 7}) {
    for (int spinners : {0, 3, 8, 11}) {
      const int warps = 8 + spinners;
      ubench<<<1, warps * 32, 0>>>(mode, iters, d_cycles, d_sink, 8);
      ubench<<<1, warps * 32, 0>>>(mode, iters, d_cycles, d_sink, 8);
      cudaDeviceSynchronize();
      std::vector<long long> h(32);
      cudaMemcpy(h.data(), d_cycles, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < 8; ++w) mx = h[w] > mx ? h[w] : mx;
      const double per_iter = (double)mx / iters;
      const double elems = 8 * 32.0 * (mode == 7 ? 112 : 128);
      printf("== mode %d, 8 workers + %2d mbarrier spinners: %9.1f cyc/iter  %7.2f exp/cyc/SM\n", mode, spinners, per_iter,
             elems / per_iter);
    }
  }
  // co-runners: 8 workers in the exp loop while 8 other warps run the max pass (LDTM + FMNMX) or a second exp loop
  for (int comode : {6, 7}) {
    ubench<<<1, 16 * 32, 0>>>(7, iters, d_cycles, d_sink, 8, comode);
    ubench<<<1, 16 * 32, 0>>>(7, iters, d_cycles, d_sink, 8, comode);
    cudaDeviceSynchronize();
    std::vector<long long> h(32);
    cudaMemcpy(h.data(), d_cycles, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < 8; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("== mode 7, 8 workers + 8 co-runners in mode %d: %9.1f cyc/iter  %7.2f exp/cyc/SM (workers only)\n", comode,
           (double)mx / iters, 8 * 32.0 * 112 / ((double)mx / iters));
  }
  return 0;
}
