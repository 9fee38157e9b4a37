// tma_probe.cu — where does a 4-D TMA box land in shared memory when its inner dimension (32 B) is narrower than the
// 128-byte swizzle span and the global strides are not monotonic?  (The patch view of an NCHW image: dims (px, y, gx, bc)
// with strides (1, W, patch, H*W).)  Fills one uint16 "image" with its own linear index, loads one box, dumps smem.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/build/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t* out, int c1, int c3, int nbytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xEEEE;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), s = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nbytes));
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                     "r"(s), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(0), "r"(c1), "r"(0), "r"(c3) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  const int W = 224, H = 224, C = 3, B = 2, ps = 16, gw = W / ps, gwp = 16;
  std::vector<uint16_t> img((size_t)B * C * H * W);
  for (int bc = 0; bc < B * C; ++bc)
    for (int i = 0; i < H * W; ++i) img[(size_t)bc * H * W + i] = (uint16_t)(i & 0xffff);   // y * W + x (< 65536)
  uint16_t *d_img, *d_out;
  cudaMalloc(&d_img, img.size() * 2);
  cudaMalloc(&d_out, 16384);
  cudaMemcpy(d_img, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<EncodeTiledFn>(fp);
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)ps, (cuuint64_t)H, (cuuint64_t)gw, (cuuint64_t)(B * C)};
    cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)ps * 2, (cuuint64_t)H * W * 2};
    cuuint32_t box[4] = {(cuuint32_t)ps, (cuuint32_t)(64 / ps), (cuuint32_t)gwp, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %s: encode -> %d\n", mode == 0 ? "SWIZZLE_128B" : "SWIZZLE_NONE", (int)r);
    if (r != CUDA_SUCCESS) continue;
    const int gy = 3, py0 = 4;
    probe<<<1, 128, 16384>>>(tm, d_out, gy * ps + py0, 1, ps * (64 / ps) * gwp * 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> out(8192);
    cudaMemcpy(out.data(), d_out, 16384, cudaMemcpyDeviceToHost);
    // expected dense layout: element (gx', py, px) at u16 index gx' * 64 + py * 16 + px (before swizzle) with value (gy*16+py0+py) * W + gx' * 16 + px
    int last = -1;
    for (int i = 0; i < 8192; ++i) if (out[i] != 0xEEEE) last = i;
    printf("  last written u16 index: %d (dense box = 1024 elements)\n", last);
    for (int row = 0; row < 20; ++row) {   // print 128-byte rows: decode each 16-byte chunk's first element as (y, x)
      printf("  smem row %2d:", row);
      for (int ch = 0; ch < 8; ++ch) {
        const uint16_t v = out[row * 64 + ch * 8];
        if (v == 0xEEEE) printf("  [----,---]"); else printf("  [y%3d,x%3d]", v / W, v % W);
      }
      printf("\n");
    }
  }
  return 0;
}
