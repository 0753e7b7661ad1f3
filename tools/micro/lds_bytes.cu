// Microbenchmark: does LDS.U8 merge lanes that read different bytes of one
// 32-bit word?  nvcc -arch=sm_100a -O3 -o lds_bytes lds_bytes.cu && ./lds_bytes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(const uint32_t *offs, int pattern, unsigned long long *out, uint32_t *sink) {
  __shared__ unsigned char buf[16384];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) buf[i] = (unsigned char)(i * 7);
  __syncthreads();
  uint32_t a = (uint32_t)__cvta_generic_to_shared(buf) + offs[pattern * 32 + (threadIdx.x & 31)];
  uint32_t acc = 0;
  __syncthreads();
  unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      uint32_t r;
      asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(r) : "r"(a + ((u & 3) << 10)));
      acc += r;
    }
  }
  buf[8192 + threadIdx.x] = (unsigned char)acc;  // waits for every load of this thread
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) out[pattern] = t1 - t0;
  sink[threadIdx.x] = acc;
}

int main() {
  const int NP = 8;
  uint32_t h[NP * 32];
  const char *names[NP] = {"32 lanes, 32 distinct words (no conflict)", "all lanes one byte (broadcast)",
                           "4 bytes of ONE word", "bytes 0..31 (8 words x 4 bytes)",
                           "2-way bank conflict (words w, w+32)", "4-way bank conflict",
                           "same word, bytes lane&1", "8 words x 4 bytes, words in distinct banks, +2-way conflict"};
  for (int l = 0; l < 32; ++l) {
    h[0 * 32 + l] = l * 4;
    h[1 * 32 + l] = 5;
    h[2 * 32 + l] = l & 3;
    h[3 * 32 + l] = l;
    h[4 * 32 + l] = (l & 15) * 4 + (l >> 4) * 128;
    h[5 * 32 + l] = (l & 7) * 4 + (l >> 3) * 128;
    h[6 * 32 + l] = l & 1;
    h[7 * 32 + l] = (l & 15) + (l >> 4) * 128;
  }
  uint32_t *d; unsigned long long *o; uint32_t *sink;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, NP * 8); cudaMalloc(&sink, 4096);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int p = 0; p < NP; ++p) probe<<<1, 1024>>>(d, p, o, sink);
  unsigned long long r[NP];
  cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
  for (int p = 0; p < NP; ++p)
    printf("%-70s %6.2f cycles per warp-LDS (32 warps)\n", names[p], (double)r[p] / (256.0 * 16 * 32));
  return 0;
}
