// Probe: does a 5-D tiled TMA box (ix, iy, px, py, image) with SWIZZLE_128B land as the K-major 128-byte-row layout
// tcgen05.mma expects (row = patch, 64 k-values = 4 pixel rows x 16 pixels, 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4))?
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_im2col_probe tma_im2col_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t* out, int c1, int c2, int c3, int img) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(16384));
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar_a), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(img) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.b32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a));
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int H = 128, W = 512, P = 16, NIMG = 2, gw = W / P, gh = H / P;
  std::vector<uint16_t> h((size_t)NIMG * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i % 65521);   // unique-ish tags
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 16384);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[5] = {(cuuint64_t)P, (cuuint64_t)P, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)NIMG};
  cuuint64_t st[4] = {(cuuint64_t)W * 2, (cuuint64_t)P * 2, (cuuint64_t)P * W * 2, (cuuint64_t)H * W * 2};
  cuuint32_t box[5] = {(cuuint32_t)P, 64 / P, (cuuint32_t)gw, 128 / gw, 1}, es[5] = {1, 1, 1, 1, 1};
  if (mode == 1) {   // monotonic strides: (ix, px, iy, py, img)
    dims[1] = gw; dims[2] = P; st[0] = P * 2; st[1] = W * 2; box[1] = gw; box[2] = 64 / P;
  }
  if (mode == 5) {   // (ix, px, py, iy, img), SWIZZLE_32B: smem = [iy][patch][16 ix], 32-byte rows
    dims[1] = gw; dims[2] = gh; dims[3] = P;
    st[0] = P * 2; st[1] = (cuuint64_t)P * W * 2; st[2] = W * 2;
    box[1] = gw; box[2] = 128 / gw; box[3] = 64 / P;
  }
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      mode == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : (mode == 5 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                                      mode == 4 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  const int iy0 = 4, py0 = 4, img = 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 17408);
  if (mode == 5) probe<<<1, 128, 17408>>>(tm, o, 0, py0, iy0, img);
  else probe<<<1, 128, 17408>>>(tm, o, iy0, 0, py0, img);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint16_t> res(8192);
  cudaMemcpy(res.data(), o, 16384, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r0 = 0; r0 < 128; ++r0)          // row = patch (py, px) of the box
    for (int k = 0; k < 64; ++k) {          // k = (iy, ix) inside the 64-wide K block
      const int py = py0 + r0 / gw, px = r0 % gw, iy = iy0 + k / P, ix = k % P;
      const size_t src = ((size_t)img * H + (size_t)py * P + iy) * W + (size_t)px * P + ix;
      const int chunk = k / 8, within = k % 8;
      int pos = r0 * 64 + ((chunk ^ (r0 & 7)) * 8) + within;
      if (mode == 5) {   // [iy][row][16 ix], 16-byte chunk c (0/1) of row r stored at chunk c ^ ((r >> 2) & 1)
        const int iyl = k / P, c2 = (k % P) / 8;
        pos = iyl * 128 * 16 + r0 * 16 + ((c2 ^ ((r0 >> 2) & 1)) * 8) + within;
      }
      if (res[pos] != h[src]) { if (bad < 5) printf("mismatch row %d k %d: got %u want %u\n", r0, k, res[pos], h[src]); ++bad; }
    }
  printf("swizzled K-major layout check: %d mismatches of 8192\n", bad);
  return 0;
}
