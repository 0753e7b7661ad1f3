// Microbenchmark: how many small bulk copies per second can one SM issue?
// Every lane of the producer warps brings ONE line-sized box into shared memory,
//   mode 0: cp.async.bulk (1-D, source and destination 16-byte aligned),
//   mode 1: cp.async.bulk.tensor.1d through a u8 tensor map, box starting at an
//           arbitrary byte (the hardware re-aligns the line to the row's start).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_lines tma_lines.cu && ./tma_lines
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

template <int BOX>
__global__ void __launch_bounds__(1024, 1)
probe(const __grid_constant__ CUtensorMap tmap, const uint8_t *text, uint64_t textBytes, int mode, int slots,
      int iters, uint32_t lineStep, uint32_t *sink, uint32_t alignMask) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nWarps = blockDim.x >> 5;
  // per warp: `slots` slots of 32 rows of BOX bytes; barriers behind the slots
  const uint32_t slotBytes = 32 * BOX;
  const uint32_t ringA = smem_addr(smem) + warp * slots * slotBytes;
  const uint32_t barA = smem_addr(smem) + nWarps * slots * slotBytes + warp * slots * 8;
  if (lane == 0)
    for (int s = 0; s < slots; ++s)
      mbar_init(barA + 8 * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // the lines of this warp: consecutive in the text, lineStep bytes apart, odd start
  uint64_t line = (uint64_t)(blockIdx.x * nWarps + warp) * iters * 32 + lane;
  for (int it = 0; it < iters; ++it, line += 32) {
    const int s = it % slots;
    if (it >= slots)
      mbar_wait(barA + 8 * s, ((it / slots) - 1) & 1);
    __syncwarp();
    if (lane == 0)
      mbar_arrive_tx(barA + 8 * s, slotBytes);
    __syncwarp();
    const uint64_t start = (line * lineStep + 3) % (textBytes - 2 * BOX);
    const uint32_t dst = ringA + s * slotBytes + lane * BOX;
    if (mode == 0) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(text + (start & ~15ull)), "r"(BOX), "r"(barA + 8 * s) : "memory");
    }
    else if (mode == 1) {
      asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
                   ::"r"(dst), "l"(&tmap), "r"((int32_t)(start & ~(uint64_t)alignMask)), "r"(barA + 8 * s) : "memory");
    }
    else {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(dst), "l"(&tmap), "r"((int32_t)(start & 0xfffff)), "r"((int32_t)(start >> 20)), "r"(barA + 8 * s) : "memory");
    }
  }
  for (int it = (iters > slots ? iters - slots : 0); it < iters; ++it)
    mbar_wait(barA + 8 * (it % slots), (it / slots) & 1);
  // check what landed: row `lane` of the last slot starts with the line's first byte
  if (sink) {
    const int it = iters - 1;
    const uint64_t l = (uint64_t)(blockIdx.x * nWarps + warp) * iters * 32 + lane + (uint64_t)it * 32;
    const uint64_t start = (l * lineStep + 3) % (textBytes - 2 * BOX);
    const unsigned char *row = smem + warp * slots * slotBytes + (it % slots) * slotBytes + lane * BOX;
    const uint8_t want = (mode == 0) ? text[start & ~15ull] : text[start & ~(uint64_t)alignMask];  // (mode 2: a box that crosses a row end is cut there)
    const uint8_t want2 = (mode == 0) ? text[(start & ~15ull) + BOX - 1] : text[(start & ~(uint64_t)alignMask) + BOX - 1];
    if ((row[0] != want) || (row[BOX - 1] != want2))
      atomicAdd(sink, 1u);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BOX>
void run(EncodeFn encode, uint8_t *text, uint64_t bytes, uint32_t *sink, int warps, int slots, int mode, uint32_t lineStep) {
  CUtensorMap tmap;
  cuuint64_t dim[2] = {bytes, 1};
  cuuint64_t stride[2] = {0, 0};
  cuuint32_t box[2] = {BOX, 1};
  cuuint32_t estr[2] = {1, 1};
  if (mode == 2) {
    dim[0] = 1u << 20;
    dim[1] = bytes >> 20;
    stride[0] = 1u << 20;
  }
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, mode == 2 ? 2 : 1, text, dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    return;
  }
  const int iters = 2000;
  const size_t smemBytes = (size_t)warps * slots * 32 * BOX + warps * slots * 8 + 128;
  if (smemBytes > 227 * 1024) {
    printf("box %d warps %d slots %d: too much shared memory\n", BOX, warps, slots);
    return;
  }
  cudaFuncSetAttribute(probe<BOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  cudaMemset(sink, 0, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<BOX><<<148, warps * 32, smemBytes>>>(tmap, text, bytes, mode, slots, iters, lineStep, sink, getenv("ALIGN") ? (uint32_t)atoi(getenv("ALIGN")) : 0u);
  cudaEventRecord(e0);
  probe<BOX><<<148, warps * 32, smemBytes>>>(tmap, text, bytes, mode, slots, iters, lineStep, sink, getenv("ALIGN") ? (uint32_t)atoi(getenv("ALIGN")) : 0u);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  uint32_t bad = 0;
  cudaMemcpy(&bad, sink, 4, cudaMemcpyDeviceToHost);
  const double ops = 148.0 * warps * 32 * iters;
  printf("mode %d box %3d warps %2d slots %2d step %4u: %8.3f ms  %7.2f Gops/s  %6.1f cycles/op/SM @1.965GHz  %7.1f GB/s  bad %u  %s\n",
         mode, BOX, warps, slots, lineStep, ms, ops / ms / 1e6, ms * 1e-3 * 1.965e9 / (ops / 148), ops * BOX / ms / 1e6, bad,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char **argv) {
  const uint64_t bytes = 1ull << 30;
  uint8_t *text;
  uint32_t *sink;
  cudaMalloc(&text, bytes);
  cudaMalloc(&sink, 4);
  {
    uint8_t *h = (uint8_t *)malloc(bytes);
    for (uint64_t i = 0; i < bytes; ++i)
      h[i] = (uint8_t)((i * 2654435761u) >> 13);
    cudaMemcpy(text, h, bytes, cudaMemcpyHostToDevice);
    free(h);
  }
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q);
  if (!encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  if (argc < 6) {
    printf("usage: tma_lines MODE BOX WARPS SLOTS STEP\n");
    return 1;
  }
  const int mode = atoi(argv[1]), box = atoi(argv[2]), warps = atoi(argv[3]), slots = atoi(argv[4]);
  const uint32_t step = (uint32_t)atoi(argv[5]);
  if (box == 64) run<64>(encode, text, bytes, sink, warps, slots, mode, step);
  else if (box == 128) run<128>(encode, text, bytes, sink, warps, slots, mode, step);
  else run<256>(encode, text, bytes, sink, warps, slots, mode, step);
  return 0;
}
