/* stream.h - line index, match compaction and the long-stream matcher
   (implemented in stream.cu). */

#pragma once

#include <cstdint>
#include <string>
#include <cuda_runtime.h>

#include "../../include/red_b200.h"
#include "device_program.h"
#include "program.h"

namespace red_b200 {

int split_lines(const uint8_t *dText, uint64_t len, uint8_t delim,
                uint64_t capacity, uint64_t *dStarts, uint32_t *dLens,
                uint64_t *nLinesOut, cudaStream_t stream, std::string &why);

int compact_matches(const int32_t *dResults, uint64_t n, uint64_t *dIndices,
                    uint64_t *dCount, cudaStream_t stream, std::string &why);

int match_stream(const ProgramDev &pd, bool smemTable, const DeviceImage &img,
                 const LaunchGeometry &geo, const uint8_t *dText, uint64_t len,
                 int style, red_b200_outcome *out, cudaStream_t stream,
                 int *launches, std::string &why);

} // namespace red_b200
