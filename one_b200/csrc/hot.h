/* hot.h - the staged check/match batch kernel (kernels_hot.cu): what the host
   side hands it and its launch entry points. */

#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "device_program.h"

namespace red_b200 {

// device copy of a HotImage (program.h); all pointers are device pointers
struct HotDev {
  const uint8_t *table = nullptr;     // (nHot + 1) rows of `stride` bytes
  uint32_t stride = 0;                // columns * entryBytes + 4 (program.h)
  uint32_t entryBytes = 1;            // 1, or 2 (the long-stream matcher's image: more than 255 states)
  const uint32_t *codeToV = nullptr;  // [max(256, nHot + 1)]
  const uint8_t *codeOf = nullptr;    // [stateCount] entries of entryBytes, by renumbered state index
  uint32_t tableBytes = 0;            // multiple of 16
  uint32_t nHot = 0;                  // the sink code
  uint32_t cAccept = 0, cAbs = 0, cPure = 0;
  uint32_t initCode = 0;
  uint32_t wideTable = 0;             // full table: 1 = u32 entries (global mode), 0 = u16
  uint32_t coverage = 1000;           // per mille of the profiled state entries that the hot set holds
  uint32_t acceptSticks = 0;          // accepting codes loop to themselves (the image of style Instant)
};

struct HotTuning {
  int consumerWarps = 0;   // 0 = pick (all the warps the ring can feed)
  int producers = 0;       // 0 = pick (4)
  int defer = 0;           // careful passes at the end of a tile: 0 = pick, 1 = yes, -1 = no
  bool force = false;      // take this kernel even when the hot set leaves out states the profile saw
  int slots = 0;           // 0 = pick (as many as fit)
  int batchLines = 0;      // 0 = pick (256); a multiple of 32, <= 256
  int slotBytes = 0;       // 0 = pick
  unsigned long long *debug = nullptr;  // device array of 16 cycle counters (tools/hot_sweep.py), or null
};

/* Styles Instant, Last and Full of verbs check/match without a leader check;
   returns cudaErrorNotSupported (nothing launched) for anything else or when
   the image does not fit, and the caller takes the other kernels.
   `counter` is a zeroable device word; avgLineBytes = 0: unknown. */
cudaError_t launch_batch_hot(const ProgramDev &pd, const HotDev &hd, const BatchArgs &args,
                             int style, const LaunchGeometry &geo, const HotTuning &tune,
                             uint32_t avgLineBytes, unsigned long long *counter,
                             cudaStream_t stream, int *launches);

/* Walks a sample of the batch's lines (every `stride`-th) to the end on the
   full table and counts how often every state is ENTERED (a lane that leaves
   the hot set pays per excursion, not per byte): hist[index] up to and
   including the first accepting state of a line - what an Instant walk sees -
   and hist[nStates + index] behind it.  hist: zeroed device array of
   2 * nStates u64; lenSum (two device u64, zeroed): the sampled lines' total
   length and how many of their bytes are >= 128. */
cudaError_t launch_profile(const ProgramDev &pd, uint32_t wideTable, const BatchArgs &args,
                           uint64_t stride, uint32_t nStates, unsigned long long *hist,
                           unsigned long long *lenSum, cudaStream_t stream);

} // namespace red_b200
