// ubench_alu.cu — per-SM throughput of the CUDA-core instructions the attention softmax loops are made of
// (sm_100a): FFMA, FADD, FMNMX (2- and 3-input), F2FP bf16x2 pack, MUFU.EX2, and an FMA-pipe polynomial exp2.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/ubench_alu tools/ubench_alu.cu
#include <cstdio>
#include <vector>

#include "../vision_transformers_torch_xla_b200/csrc/vitk_common.cuh"

using namespace vitk;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// exp2 on the FMA pipe: round-to-nearest split + degree-3 polynomial on [-0.5, 0.5] + exponent insertion
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float fi = t - 12582912.0f;
  const float f = x - fi;                   // [-0.5, 0.5]
  float p = fmaf(f, 0.0555041086f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

constexpr int U = 16;  // independent chains per thread

__global__ void __launch_bounds__(1024, 1) ubench(int mode, int iters, long long* cycles, float* sink) {
  const int warp = threadIdx.x >> 5;
  float x[U];
#pragma unroll
  for (int i = 0; i < U; ++i) x[i] = 0.001f * (float)(i + threadIdx.x);
  float y = 0.37f + 0.001f * threadIdx.x, z = 1.0001f;
  uint32_t pk[U / 2];
#pragma unroll
  for (int i = 0; i < U / 2; ++i) pk[i] = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = fmaf(x[i], z, y);
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = x[i] + y;
    } else if (mode == 2) {
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = fmaxf(x[i], y + (float)it);
    } else if (mode == 3) {
#pragma unroll
      for (int i = 0; i < U; i += 2) x[i] = fmax3(x[i], x[i + 1], y + (float)it);
    } else if (mode == 4) {
#pragma unroll
      for (int i = 0; i < U / 2; ++i) pk[i] ^= pack_bf16x2(x[2 * i] + (float)it, x[2 * i + 1]);
    } else if (mode == 5) {
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = ex2_approx(x[i] - 1.0f);
    } else if (mode == 6) {
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = exp2_poly(x[i] - 1.0f);
    } else if (mode == 7) {
      // 3/4 MUFU + 1/4 polynomial
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = (i & 3) == 3 ? exp2_poly(x[i] - 1.0f) : ex2_approx(x[i] - 1.0f);
    } else if (mode == 8) {
      // the softmax element: FFMA -> EX2 -> FADD, pack pairs
#pragma unroll
      for (int i = 0; i < U; i += 2) {
        const float e0 = ex2_approx(fmaf(x[i], z, -y)), e1 = ex2_approx(fmaf(x[i + 1], z, -y));
        x[i] = e0 + x[i];
        x[i + 1] = e1 + x[i + 1];
        pk[i / 2] ^= pack_bf16x2(e0, e1);
      }
    } else if (mode == 9) {
      // the same with every 4th exp on the FMA pipe
#pragma unroll
      for (int i = 0; i < U; i += 2) {
        const float a0 = fmaf(x[i], z, -y), a1 = fmaf(x[i + 1], z, -y);
        const float e0 = ex2_approx(a0), e1 = (i & 2) ? exp2_poly(a1) : ex2_approx(a1);
        x[i] = e0 + x[i];
        x[i + 1] = e1 + x[i + 1];
        pk[i / 2] ^= pack_bf16x2(e0, e1);
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) acc += x[i];
#pragma unroll
  for (int i = 0; i < U / 2; ++i) acc += __uint_as_float(pk[i]);
  if ((threadIdx.x & 31) == 0) cycles[warp] = t1 - t0;
  if (acc == 12345.678f) sink[threadIdx.x] = acc;
}

int main() {
  long long* d_cycles;
  float* d_sink;
  cudaMalloc(&d_cycles, 32 * sizeof(long long));
  cudaMalloc(&d_sink, 1024 * sizeof(float));
  const int iters = 512;
  const char* names[] = {"FFMA", "FADD", "FMNMX", "FMNMX3 (2 elements per instruction)", "F2FP bf16x2 pack (+FADD)", "MUFU.EX2 (+FADD)",
                         "exp2 polynomial (FMA pipe)", "3/4 MUFU + 1/4 polynomial", "softmax element (FFMA, EX2, FADD, pack)",
                         "softmax element, 1/4 of exps polynomial"};
  const double per_iter_elems[] = {U, U, U, U, U, U, U, U, U, U};
  for (int mode = 0; mode < 10; ++mode) {
    printf("== mode %d: %s\n", mode, names[mode]);
    for (int warps : {4, 8, 16}) {
      ubench<<<1, warps * 32>>>(mode, iters, d_cycles, d_sink);
      ubench<<<1, warps * 32>>>(mode, iters, d_cycles, d_sink);
      if (cudaDeviceSynchronize() != cudaSuccess) return 1;
      std::vector<long long> h(32);
      cudaMemcpy(h.data(), d_cycles, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
      const double cyc = (double)mx / iters;
      printf("   warps=%2d  %8.1f cyc/iter  %7.2f elements/cyc/SM\n", warps, cyc, warps * 32.0 * per_iter_elems[mode] / cyc);
    }
  }
  return 0;
}
