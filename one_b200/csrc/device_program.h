/* device_program.h - what the kernels see of an uploaded program, and the
   launch entry points implemented in kernels.cu. */

#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace red_b200 {

// Passed to kernels by value.  All pointers are device pointers.
struct ProgramDev {
  const uint16_t *clsOffset;  // [256] byte -> class * entryBytes
  const void *table;          // u16 (smem mode) or u32 (global mode) entries
  const int32_t *results;     // [stateCount], by renumbered state index
  const uint8_t *equiv;       // [256] byte -> class (as serialized)
  const uint8_t *leader;      // [leaderLen] required prefix, as classes
  uint32_t leaderLen;
  uint32_t leaderV;           // state reached after the leader
  uint32_t tableBytes;
  uint32_t unitShift;   // smem mode: row entry index = v << unitShift
  uint32_t initV;
  uint32_t acceptV;     // acceptV <= v < pureV : accepting
  uint32_t absV;        // absV <= v < pureV   : accepting, can never be left
  uint32_t pureV;       // v >= pureV          : pure dead end
  uint32_t vPerState;
  uint32_t divMagic;    // 0, or M with (v * M) >> 32 == v / vPerState for all v
  uint32_t clsMaxOffset;  // largest value in clsOffset
  int32_t initResult;
  uint32_t byteRows;    // 1: 256-column rows indexed by the raw byte (unitShift 8)
  uint32_t hasFinalStates;  // 1: a state >= absV is reachable from the initial state
  uint32_t stateCount;  // v < stateCount * vPerState for every state of the program
  const uint8_t *resetCount;    // [256], see DeviceImage::resetCount (program.h)
  const uint32_t *resetStates;  // [256 * 8]
  uint32_t hotBytes;    // global-table mode: leading bytes of the table mirrored in shared
                        // memory by the production kernels (0: none)
  /* scan/search/replace try an anchored walk at every start position; nearly
     all of them die within two bytes.  startFilter says which cannot match,
     from the bytes at the position and behind it (computed at upload):
     bit (b0 << 8 | b1) of words [0, 2048): the state behind b0 accepts, or
     the one behind b0 b1 is not a dead end; bit b0 of words [2048, 2056), for
     a line's last byte: the state behind b0 accepts.  A position whose bit is
     clear yields no match in any style. */
  const uint32_t *startFilter;
};

struct BatchArgs {
  const uint8_t *text;
  const uint64_t *starts;
  const uint32_t *lens;
  uint64_t nLines;
  int32_t *results;
  uint32_t *matchStarts;  // may be null
  uint32_t *matchEnds;    // may be null
  bool checkLeader = false;  // verb check on a program that has a leader
  const uint64_t *nLinesDev = nullptr;  // if set: the line count lives on the device
                                        // (nLines is then only an upper bound)
};

struct LaunchGeometry {
  int smCount;
  int maxSmemOptin;  // bytes
};

// tuning knobs, overridable for sweeps (bench.py --tune)
struct BatchTuning {
  int threads = 0;      // 0 = pick
  int ctasPerSm = 0;    // 0 = pick
  bool simpleWalk = false;  // per-byte decisions only (A/B baseline)
  int variant = 0;      // 0 = tile schedule (production), 3 = persistent lanes,
                        // 1 = first kernel (kernels.cu), kept as the A/B baseline
  int linesPerLane = 0; // persistent lanes: lines per lane per claimed block (0 = pick)
};

// verbs check/match (anchored walk).  smemTable selects the table mode.
cudaError_t launch_batch_anchored(const ProgramDev &pd, bool smemTable,
                                  const BatchArgs &args, int style,
                                  const LaunchGeometry &geo,
                                  const BatchTuning &tune, cudaStream_t stream,
                                  int *launches);

// verbs check/match, persistent lanes (kernels_lanes.cu): the production path.
// `counter` is a zeroable device word used to hand out blocks of lines.
cudaError_t launch_batch_lanes(const ProgramDev &pd, bool smemTable,
                               const BatchArgs &args, int style,
                               const LaunchGeometry &geo,
                               const BatchTuning &tune,
                               unsigned long long *counter,
                               cudaStream_t stream, int *launches);

// one outcome of matchAll()/collect(), positions relative to the line
struct MultiSpan {
  int32_t result;
  uint32_t start;
  uint32_t end;
};

// per line: all matches (kernels_multi.cu); which = 0 matchAll, 1 collect
cudaError_t launch_batch_multi(const ProgramDev &pd, bool smemTable, const uint8_t *text,
                               const uint64_t *starts, const uint32_t *lens, uint64_t nLines,
                               int which, uint32_t maxPerLine, uint32_t *counts,
                               MultiSpan *spans, const LaunchGeometry &geo,
                               cudaStream_t stream, int *launches);

// replace() per line in two passes (kernels_multi.cu): write = false fills
// counts/outLens, write = true writes the rewritten lines at outStarts
cudaError_t launch_batch_replace(const ProgramDev &pd, bool smemTable, bool write,
                                 const uint8_t *text, const uint64_t *starts,
                                 const uint32_t *lens, uint64_t nLines, int style,
                                 uint64_t maxRepl, const uint8_t *repl, uint32_t replLen,
                                 int delim, uint32_t *counts, uint32_t *outLens,
                                 const uint64_t *outStarts, uint8_t *outText,
                                 const LaunchGeometry &geo, cudaStream_t stream, int *launches);

// StatefulMatcher::advance over one segment per line, state carried in states[]
cudaError_t launch_batch_advance(const ProgramDev &pd, bool smemTable, const uint8_t *text,
                                 const uint64_t *starts, const uint32_t *lens, uint64_t nLines,
                                 uint32_t *states, int32_t *results, int32_t *hitResults,
                                 uint32_t *hitEnds, const LaunchGeometry &geo,
                                 cudaStream_t stream, int *launches);

// verbs scan/search (sliding start position)
cudaError_t launch_batch_sliding(const ProgramDev &pd, bool smemTable,
                                 const BatchArgs &args, int style,
                                 bool wantPositions,
                                 const LaunchGeometry &geo,
                                 cudaStream_t stream, int *launches,
                                 unsigned long long *counter = nullptr,  // zeroable device word: enables
                                                                         // the warp-per-line kernel
                                 bool simple = false);                   // force the lane-per-line kernel

} // namespace red_b200
