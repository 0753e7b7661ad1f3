/* stream.cu - line index (split at a delimiter), order-preserving compaction
   of matching lines, and the long-stream matcher.

   Index and compaction are the same three-step shape, sized for HBM:
     1. every warp counts the hits in its own contiguous segment,
     2. one block turns the per-segment counts into exclusive offsets,
     3. every warp walks its segment again and writes each hit at
        offset + (ballot rank inside the warp).
   Text is read twice, 16 bytes per lane, fully coalesced. */

#include "stream.h"

#include <algorithm>
#include <mutex>
#include <vector>

namespace red_b200 {

namespace {

constexpr int kScanThreads = 1024;

/* nDelims delimiters give nDelims lines plus one more if the text does not
   end with a delimiter; clamped to the caller's capacity (overflow set) */
__global__ void line_count_kernel(const uint8_t *text, uint64_t len, uint8_t delim,
                                  uint64_t capacity, const uint64_t *total,
                                  LineCount *info) {
  uint64_t nd = *total;
  uint64_t n = nd + ((len && (text[len - 1] != delim)) ? 1 : 0);
  info->wanted = n;
  info->overflow = n > capacity;
  info->nLines = n > capacity ? capacity : n;
  info->nDelims = nd;
}
constexpr uint32_t kSegBytes = 16384;  // text bytes per warp segment
constexpr uint32_t kSegLines = 4096;   // results per warp segment

__device__ __forceinline__ uint32_t match_mask16(const uint4 &q, uint32_t d4) {
  // bit j set <=> byte j of the 16-byte block equals the delimiter
  uint32_t m = 0;
  uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t eq = __vcmpeq4(w[i], d4);  // 0xff per equal byte
    // gather the top bit of each byte into 4 adjacent bits
    uint32_t bits = ((eq & 0x80808080u) * 0x00204081u) >> 28;
    m |= bits << (4 * i);
  }
  return m;
}

__global__ void count_delims_kernel(const uint8_t *text, uint64_t len,
                                    uint32_t d4, uint32_t *segCounts,
                                    uint64_t nSegs) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nWarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t seg = warp; seg < nSegs; seg += nWarps) {
    const uint64_t base = seg * kSegBytes;
    const uint64_t end = min(len, base + kSegBytes);
    uint32_t cnt = 0;
    for (uint64_t off = base + lane * 16; off < end; off += 512) {
      uint4 q = __ldg(reinterpret_cast<const uint4 *>(text + off));
      uint32_t m = match_mask16(q, d4);
      if (off + 16 > end)
        m &= (1u << (end - off)) - 1;
      cnt += __popc(m);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0)
      segCounts[seg] = cnt;
  }
}

__global__ void count_matches_kernel(const int32_t *results, uint64_t n,
                                     uint32_t *segCounts, uint64_t nSegs,
                                     const uint64_t *nDev = nullptr) {
  if (nDev)
    n = min(n, *nDev);
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nWarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t seg = warp; seg < nSegs; seg += nWarps) {
    const uint64_t base = seg * kSegLines;
    const uint64_t end = min(n, base + kSegLines);
    uint32_t cnt = 0;
    for (uint64_t i = base + lane; i < end; i += 32)
      cnt += (__ldg(results + i) > 0);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0)
      segCounts[seg] = cnt;
  }
}

/* exclusive scan of segCounts (u32) into segOffsets (u64); total to *total.
   One block; each thread owns a contiguous run so the pass is two sweeps. */
__global__ void __launch_bounds__(kScanThreads)
scan_counts_kernel(const uint32_t *segCounts, uint64_t *segOffsets,
                   uint64_t nSegs, uint64_t *total) {
  __shared__ uint64_t partial[kScanThreads];
  const uint64_t per = (nSegs + kScanThreads - 1) / kScanThreads;
  const uint64_t lo = min(nSegs, per * threadIdx.x);
  const uint64_t hi = min(nSegs, lo + per);
  uint64_t sum = 0;
  for (uint64_t i = lo; i < hi; ++i)
    sum += segCounts[i];
  partial[threadIdx.x] = sum;
  __syncthreads();
  // Hillis-Steele over 1024 partials
  for (int d = 1; d < kScanThreads; d <<= 1) {
    uint64_t add = (threadIdx.x >= d) ? partial[threadIdx.x - d] : 0;
    __syncthreads();
    partial[threadIdx.x] += add;
    __syncthreads();
  }
  uint64_t run = partial[threadIdx.x] - sum;  // exclusive
  for (uint64_t i = lo; i < hi; ++i) {
    segOffsets[i] = run;
    run += segCounts[i];
  }
  if (threadIdx.x == kScanThreads - 1)
    *total = partial[kScanThreads - 1];
}

/* Delimiter number r (0-based) at position d ends line r and starts line r+1:
     lens[r]     = low 32 bits of d      (fixed up to a length afterwards)
     starts[r+1] = d + 1                 (if that line exists)            */
__global__ void write_lines_kernel(const uint8_t *text, uint64_t len,
                                   uint32_t d4, const uint64_t *segOffsets,
                                   uint64_t nSegs, const LineCount *info,
                                   uint64_t *starts, uint32_t *lens) {
  const uint64_t nLines = info->nLines;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nWarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t seg = warp; seg < nSegs; seg += nWarps) {
    const uint64_t base = seg * kSegBytes;
    const uint64_t end = min(len, base + kSegBytes);
    uint64_t rank = segOffsets[seg];
    for (uint64_t off0 = base; off0 < end; off0 += 512) {
      const uint64_t off = off0 + lane * 16;
      uint32_t m = 0;
      if (off < end) {
        uint4 q = __ldg(reinterpret_cast<const uint4 *>(text + off));
        m = match_mask16(q, d4);
        if (off + 16 > end)
          m &= (1u << (end - off)) - 1;
      }
      uint32_t c = __popc(m);
      // exclusive prefix of c across the warp
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
          incl += up;
      }
      uint64_t r = rank + (incl - c);
      while (m) {
        uint32_t j = __ffs(m) - 1;
        m &= m - 1;
        uint64_t d = off + j;
        if (r < nLines)
          lens[r] = static_cast<uint32_t>(d);
        if (r + 1 < nLines)
          starts[r + 1] = d + 1;
        ++r;
      }
      rank += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

__global__ void finish_lines_kernel(uint64_t len, const LineCount *info,
                                    uint64_t *starts, uint32_t *lens) {
  const uint64_t nLines = info->nLines, nDelims = info->nDelims;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       i < nLines; i += stride) {
    uint64_t s = (i == 0) ? 0 : starts[i];
    if (i == 0)
      starts[0] = 0;
    uint32_t e = (i < nDelims) ? lens[i] : static_cast<uint32_t>(len);
    lens[i] = e - static_cast<uint32_t>(s);
  }
}

__global__ void write_matches_kernel(const int32_t *results, uint64_t n,
                                     const uint64_t *segOffsets, uint64_t nSegs,
                                     uint64_t *indices, const uint64_t *nDev = nullptr) {
  if (nDev)
    n = min(n, *nDev);
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nWarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t seg = warp; seg < nSegs; seg += nWarps) {
    const uint64_t base = seg * kSegLines;
    const uint64_t end = min(n, base + kSegLines);
    uint64_t rank = segOffsets[seg];
    for (uint64_t i0 = base; i0 < end; i0 += 32) {
      const uint64_t i = i0 + lane;
      bool hit = (i < end) && (__ldg(results + i) > 0);
      uint32_t ballot = __ballot_sync(0xffffffffu, hit);
      if (hit)
        indices[rank + __popc(ballot & ((1u << lane) - 1))] = i;
      rank += __popc(ballot);
    }
  }
}

/* exclusive scan of n u32 items (each plus `extra`) into u64 offsets, in
   blocks of 1024 items: block sums, one-block scan of the sums, re-scan. */
constexpr int kScanBlockItems = 1024;

__global__ void __launch_bounds__(256)
block_sums_kernel(const uint32_t *items, uint64_t n, uint32_t extra, uint64_t *blockSums) {
  __shared__ unsigned long long warpSums[8];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanBlockItems;
  unsigned long long sum = 0;
  for (int k = 0; k < 4; ++k) {
    const uint64_t i = base + k * 256 + threadIdx.x;
    if (i < n)
      sum += static_cast<unsigned long long>(items[i]) + extra;
  }
  for (int d = 16; d > 0; d >>= 1)
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0)
    warpSums[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w)
      t += warpSums[w];
    blockSums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(uint64_t *blockSums, uint64_t nBlocks, uint64_t *total) {
  __shared__ uint64_t partial[kScanThreads];
  const uint64_t per = (nBlocks + kScanThreads - 1) / kScanThreads;
  const uint64_t lo = min(nBlocks, per * threadIdx.x);
  const uint64_t hi = min(nBlocks, lo + per);
  uint64_t sum = 0;
  for (uint64_t i = lo; i < hi; ++i)
    sum += blockSums[i];
  partial[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < kScanThreads; d <<= 1) {
    uint64_t add = (threadIdx.x >= d) ? partial[threadIdx.x - d] : 0;
    __syncthreads();
    partial[threadIdx.x] += add;
    __syncthreads();
  }
  uint64_t run = partial[threadIdx.x] - sum;
  for (uint64_t i = lo; i < hi; ++i) {
    const uint64_t v = blockSums[i];
    blockSums[i] = run;
    run += v;
  }
  if (threadIdx.x == kScanThreads - 1)
    *total = partial[kScanThreads - 1];
}

__global__ void __launch_bounds__(256)
block_offsets_kernel(const uint32_t *items, uint64_t n, uint32_t extra,
                     const uint64_t *blockSums, uint64_t *offsets) {
  __shared__ unsigned long long warpSums[8];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanBlockItems;
  // thread t owns items base + 4t .. base + 4t + 3
  unsigned long long v[4], sum = 0;
  for (int k = 0; k < 4; ++k) {
    const uint64_t i = base + threadIdx.x * 4ull + k;
    v[k] = (i < n) ? static_cast<unsigned long long>(items[i]) + extra : 0;
    sum += v[k];
  }
  unsigned long long incl = sum;
  const unsigned lane = threadIdx.x & 31;
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d)
      incl += up;
  }
  if (lane == 31)
    warpSums[threadIdx.x >> 5] = incl;
  __syncthreads();
  unsigned long long run = blockSums[blockIdx.x] + incl - sum;
  for (unsigned w = 0; w < (threadIdx.x >> 5); ++w)
    run += warpSums[w];
  for (int k = 0; k < 4; ++k) {
    const uint64_t i = base + threadIdx.x * 4ull + k;
    if (i < n)
      offsets[i] = run;
    run += v[k];
  }
}

int grid_for(uint64_t nSegs, int threads) {
  uint64_t warpsPerBlock = threads / 32;
  uint64_t blocks = (nSegs + warpsPerBlock - 1) / warpsPerBlock;
  return static_cast<int>(std::min<uint64_t>(blocks, 148 * 8));
}

int cuda_why(cudaError_t e, const char *what, std::string &why) {
  why = std::string(what) + ": " + cudaGetErrorString(e);
  return RED_B200_ECUDA;
}

#define TRY(expr)                                  \
  do {                                             \
    cudaError_t e_ = (expr);                       \
    if (e_ != cudaSuccess)                         \
      return cuda_why(e_, #expr, why);             \
  } while (0)

} // anonymous

int split_lines_async(const uint8_t *dText, uint64_t len, uint8_t delim,
                      uint64_t capacity, uint64_t *dStarts, uint32_t *dLens,
                      LineCount *dInfo, SplitScratch &sc, cudaStream_t stream,
                      std::string &why) {
  if (reinterpret_cast<uintptr_t>(dText) & 15) {
    why = "text must be 16-byte aligned";
    return RED_B200_EAPI;
  }
  const uint64_t nSegs = (len + kSegBytes - 1) / kSegBytes;
  if (nSegs > sc.capSegs) {
    why = "internal: split scratch too small";
    return RED_B200_EINTERNAL;
  }
  const uint32_t d4 = delim * 0x01010101u;
  const int threads = 256;
  const int grid = std::max(1, grid_for(nSegs, threads));
  if (len)
    count_delims_kernel<<<grid, threads, 0, stream>>>(dText, len, d4, sc.counts, nSegs);
  scan_counts_kernel<<<1, kScanThreads, 0, stream>>>(sc.counts, sc.offsets, nSegs, sc.total);
  line_count_kernel<<<1, 1, 0, stream>>>(dText, len, delim, capacity, sc.total, dInfo);
  if (len) {
    write_lines_kernel<<<grid, threads, 0, stream>>>(dText, len, d4, sc.offsets, nSegs,
                                                     dInfo, dStarts, dLens);
    finish_lines_kernel<<<148 * 4, 256, 0, stream>>>(len, dInfo, dStarts, dLens);
  }
  TRY(cudaGetLastError());
  return RED_B200_OK;
}

cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t stream) {
  static std::mutex lock;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  if ((dev < 0) || (dev >= 64))
    return cudaErrorInvalidDevice;
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> g(lock);
    if (!pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      if ((e = cudaMemPoolCreate(&pools[dev], &props)) != cudaSuccess)
        return e;
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool = pools[dev];
  }
  return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, stream);
}

void scratch_free(void *p, cudaStream_t stream) {
  if (p)
    cudaFreeAsync(p, stream);
}

cudaError_t SplitScratch::allocAsync(uint64_t maxTextBytes, cudaStream_t s) {
  capSegs = (maxTextBytes + kSegBytes - 1) / kSegBytes + 1;
  cudaError_t e;
  if ((e = scratch_alloc(&counts, capSegs * 4 + 4, s)) != cudaSuccess) return e;
  if ((e = scratch_alloc(&offsets, capSegs * 8 + 8, s)) != cudaSuccess) return e;
  return scratch_alloc(&total, 8, s);
}

void SplitScratch::releaseAsync(cudaStream_t s) {
  scratch_free(counts, s);
  scratch_free(offsets, s);
  scratch_free(total, s);
  counts = nullptr;
  offsets = nullptr;
  total = nullptr;
  capSegs = 0;
}

cudaError_t SplitScratch::alloc(uint64_t maxTextBytes) {
  capSegs = (maxTextBytes + kSegBytes - 1) / kSegBytes + 1;
  cudaError_t e;
  if ((e = cudaMalloc(&counts, capSegs * 4 + 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&offsets, capSegs * 8 + 8)) != cudaSuccess) return e;
  return cudaMalloc(&total, 8);
}

void SplitScratch::release() {
  cudaFree(counts);
  cudaFree(offsets);
  cudaFree(total);
  counts = nullptr;
  offsets = nullptr;
  total = nullptr;
  capSegs = 0;
}

int split_lines(const uint8_t *dText, uint64_t len, uint8_t delim,
                uint64_t capacity, uint64_t *dStarts, uint32_t *dLens,
                uint64_t *nLinesOut, cudaStream_t stream, std::string &why) {
  *nLinesOut = 0;
  if (len == 0)
    return RED_B200_OK;
  SplitScratch sc;
  LineCount *dInfo = nullptr;
  cudaError_t e = sc.allocAsync(len, stream);
  if (e == cudaSuccess)
    e = scratch_alloc(&dInfo, sizeof(LineCount), stream);
  if (e != cudaSuccess) {
    sc.releaseAsync(stream);
    return cuda_why(e, "split_lines scratch", why);
  }
  int rc = split_lines_async(dText, len, delim, capacity, dStarts, dLens, dInfo, sc, stream, why);
  LineCount info{};
  if (rc == RED_B200_OK) {
    e = cudaMemcpyAsync(&info, dInfo, sizeof(info), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess)
      rc = cuda_why(e, "split_lines", why);
  }
  scratch_free(dInfo, stream);
  sc.releaseAsync(stream);
  if (rc != RED_B200_OK)
    return rc;
  *nLinesOut = info.overflow ? info.wanted : info.nLines;
  if (info.overflow) {
    why = "line index capacity too small";
    return RED_B200_ELIMIT;
  }
  return RED_B200_OK;
}

int exclusive_offsets(const uint32_t *dItems, uint64_t n, uint32_t extra,
                      uint64_t *dOffsets, uint64_t *dTotal, cudaStream_t stream,
                      std::string &why) {
  if (n == 0) {
    TRY(cudaMemsetAsync(dTotal, 0, 8, stream));
    return RED_B200_OK;
  }
  const uint64_t nBlocks = (n + kScanBlockItems - 1) / kScanBlockItems;
  uint64_t *blockSums = nullptr;
  TRY(scratch_alloc(&blockSums, nBlocks * 8, stream));
  block_sums_kernel<<<static_cast<unsigned>(nBlocks), 256, 0, stream>>>(dItems, n, extra, blockSums);
  scan_block_sums_kernel<<<1, kScanThreads, 0, stream>>>(blockSums, nBlocks, dTotal);
  block_offsets_kernel<<<static_cast<unsigned>(nBlocks), 256, 0, stream>>>(dItems, n, extra, blockSums, dOffsets);
  cudaError_t e = cudaGetLastError();
  scratch_free(blockSums, stream);
  if (e != cudaSuccess)
    return cuda_why(e, "exclusive_offsets", why);
  return RED_B200_OK;
}

int compact_matches(const int32_t *dResults, uint64_t n, uint64_t *dIndices,
                    uint64_t *dCount, cudaStream_t stream, std::string &why) {
  if (n == 0) {
    TRY(cudaMemsetAsync(dCount, 0, 8, stream));
    return RED_B200_OK;
  }
  const uint64_t nSegs = (n + kSegLines - 1) / kSegLines;
  uint32_t *counts = nullptr;
  uint64_t *offsets = nullptr;
  TRY(scratch_alloc(&counts, nSegs * 4 + 4, stream));
  TRY(scratch_alloc(&offsets, nSegs * 8 + 8, stream));
  const int threads = 256;
  const int grid = grid_for(nSegs, threads);
  count_matches_kernel<<<grid, threads, 0, stream>>>(dResults, n, counts, nSegs);
  scan_counts_kernel<<<1, kScanThreads, 0, stream>>>(counts, offsets, nSegs, dCount);
  write_matches_kernel<<<grid, threads, 0, stream>>>(dResults, n, offsets, nSegs, dIndices);
  TRY(cudaGetLastError());
  TRY(cudaFreeAsync(counts, stream));
  TRY(cudaFreeAsync(offsets, stream));
  return RED_B200_OK;
}

int compact_matches_n(const int32_t *dResults, uint64_t capLines, const uint64_t *dN,
                      uint64_t *dIndices, uint64_t *dCount, SplitScratch &sc,
                      cudaStream_t stream, std::string &why) {
  const uint64_t nSegs = (capLines + kSegLines - 1) / kSegLines;
  if (nSegs > sc.capSegs) {
    why = "internal: compaction scratch too small";
    return RED_B200_EINTERNAL;
  }
  const int threads = 256;
  const int grid = std::max(1, grid_for(nSegs, threads));
  count_matches_kernel<<<grid, threads, 0, stream>>>(dResults, capLines, sc.counts, nSegs, dN);
  scan_counts_kernel<<<1, kScanThreads, 0, stream>>>(sc.counts, sc.offsets, nSegs, dCount);
  write_matches_kernel<<<grid, threads, 0, stream>>>(dResults, capLines, sc.offsets, nSegs, dIndices, dN);
  TRY(cudaGetLastError());
  return RED_B200_OK;
}

} // namespace red_b200
