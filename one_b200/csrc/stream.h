/* stream.h - line index, match compaction and the long-stream matcher
   (implemented in stream.cu). */

#pragma once

#include <cstdint>
#include <string>
#include <cuda_runtime.h>

#include "../../include/red_b200.h"
#include "device_program.h"
#include "hot.h"
#include "program.h"

namespace red_b200 {

struct LineCount {
  uint64_t nLines;    // lines indexed (clamped to the capacity)
  uint64_t nDelims;
  uint64_t wanted;    // lines the text really has
  uint64_t overflow;  // wanted > capacity
};

/* Scratch for one call, from a per-device stream-ordered pool that keeps its
   memory between calls (the default pool hands it back at every
   synchronization, and cudaMalloc/cudaFree cost milliseconds). */
cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t stream);
void scratch_free(void *p, cudaStream_t stream);
template <class T>
cudaError_t scratch_alloc(T **p, size_t bytes, cudaStream_t stream) {
  return scratch_alloc(reinterpret_cast<void **>(p), bytes, stream);
}

struct SplitScratch {
  uint32_t *counts = nullptr;
  uint64_t *offsets = nullptr;
  uint64_t *total = nullptr;
  uint64_t capSegs = 0;
  cudaError_t alloc(uint64_t maxTextBytes);                       // cudaMalloc (long-lived)
  void release();
  cudaError_t allocAsync(uint64_t maxTextBytes, cudaStream_t s);  // per call, from the pool
  void releaseAsync(cudaStream_t s);
};

// everything on `stream`, nothing read back: the count stays in *dInfo
int split_lines_async(const uint8_t *dText, uint64_t len, uint8_t delim,
                      uint64_t capacity, uint64_t *dStarts, uint32_t *dLens,
                      LineCount *dInfo, SplitScratch &scratch,
                      cudaStream_t stream, std::string &why);

int split_lines(const uint8_t *dText, uint64_t len, uint8_t delim,
                uint64_t capacity, uint64_t *dStarts, uint32_t *dLens,
                uint64_t *nLinesOut, cudaStream_t stream, std::string &why);

// offsets[i] = sum over j < i of (items[j] + extra); *dTotal = the whole sum
int exclusive_offsets(const uint32_t *dItems, uint64_t n, uint32_t extra,
                      uint64_t *dOffsets, uint64_t *dTotal, cudaStream_t stream,
                      std::string &why);

int compact_matches(const int32_t *dResults, uint64_t n, uint64_t *dIndices,
                    uint64_t *dCount, cudaStream_t stream, std::string &why);

// as compact_matches, but the number of lines is read from device memory
// (capLines is its upper bound) and the scratch is the caller's
int compact_matches_n(const int32_t *dResults, uint64_t capLines, const uint64_t *dN,
                      uint64_t *dIndices, uint64_t *dCount, SplitScratch &scratch,
                      cudaStream_t stream, std::string &why);

// hot: the hot image of the styles that walk on behind a match (hot.h), or null
int match_stream(const ProgramDev &pd, bool smemTable, const DeviceImage &img, const HotDev *hot,
                 const LaunchGeometry &geo, const uint8_t *dText, uint64_t len,
                 int style, red_b200_outcome *out, cudaStream_t stream,
                 int *launches, std::string &why);

} // namespace red_b200
