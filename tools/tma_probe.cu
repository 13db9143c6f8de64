// Stand-alone probe of bulk / tensor async copies on sm_100a (debugging aid, not product code).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void wait0(uint64_t* bar) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(0)
      : "memory");
}

// mode 0: 1-D bulk copy; mode 1: 2-D tensor copy
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const uint8_t* src, int mode, int x, int y, int bytes,
                      uint8_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = (uint64_t*)(smem + ((bytes + 127) / 128) * 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    if (mode == 0) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem)),
                   "l"(src), "r"(bytes), "r"(smem_u32(bar))
                   : "memory");
    } else {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
              smem_u32(smem)),
          "l"(&tmap), "r"(smem_u32(bar)), "r"(x), "r"(y)
          : "memory");
    }
  }
  __syncthreads();
  wait0(bar);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 1, esize = argc > 2 ? atoi(argv[2]) : 1, box_w = argc > 3 ? atoi(argv[3]) : 192,
      rows = argc > 4 ? atoi(argv[4]) : 182, promo = argc > 5 ? atoi(argv[5]) : 1;
  int W = 1024, H = 1280;  // elements
  std::vector<uint8_t> h((size_t)W * H * esize);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(1 + (i * 7) % 200);
  uint8_t *d, *o;
  cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  int bytes = box_w * rows * esize;
  cudaMalloc(&o, bytes);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {(cuuint64_t)W * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&tm, esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("mode %d esize %d box %dx%d promo %d: entry %d q %d encode -> %d\n", mode, esize, box_w, rows, promo, (int)ge, (int)q, (int)r);
  const unsigned char* tb = (const unsigned char*)&tm;
  for (int i = 0; i < 64; ++i) printf("%02x%s", tb[i], (i % 16 == 15) ? "\n" : " ");
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int x = argc > 6 ? atoi(argv[6]) : 96, y = argc > 7 ? atoi(argv[7]) : 200;
  probe<<<1, 256, bytes + 256>>>(tm, d, mode, x, y, bytes, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("result: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint8_t> res(bytes);
  cudaMemcpy(res.data(), o, res.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  if (mode == 0) { for (int i = 0; i < bytes; ++i) bad += res[i] != h[i]; }
  else for (int yy = 0; yy < rows; ++yy) for (int xx = 0; xx < box_w * esize; ++xx) {
    long gx = (long)x * esize + xx, gy = y + yy;
    uint8_t exp = (gx >= 0 && gx < (long)W * esize && gy >= 0 && gy < H) ? h[(size_t)gy * W * esize + gx] : 0;
    bad += res[(size_t)yy * box_w * esize + xx] != exp;
  }
  printf("mismatches %d\n", bad);
  return 0;
}
