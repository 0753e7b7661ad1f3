/* kernels.cu - sm_100a kernels that execute an uploaded RED program over a
   batch of byte strings (one lane per line).

   What is computed per line is exactly matchCore / checkCore of the reference
   (include/Matcher.h:363-495): walk the DFA from the initial state, one byte
   at a time, and stop according to the style.  How it is computed differs:

     * the table is the re-laid-out image of program.h: one step is
           v = T[(v << K) + clsOffset[byte]]
       two shared-memory loads, the first of which does not depend on v;
     * "did anything happen at this byte" is one compare, v >= limit, where
       limit = acceptV normally, and 0 once a First/Tangent walk has seen an
       accepting state (then every further state needs a decision);
     * results are looked up only when the walk ends;
     * text is read 16 bytes at a time per lane through the read-only path,
       with the next chunk requested before the current one is walked.

   Styles as template parameters: 1 instant, 2 first, 3 tangent, 4 last,
   5 full (Matcher.h:62-69).
 */

#include "device_program.h"
#include "walk.cuh"

#include <algorithm>

namespace red_b200 {

namespace {

using namespace dev;

/* One line, anchored.  p need not be aligned: the walk reads the aligned
   16-byte blocks that cover [p, p+len) and skips the bytes outside.  The
   blocks before/after the line share a 16-byte block with valid bytes, so the
   over-read stays inside the allocation granule of any CUDA allocation. */
template <int STYLE, bool WANT_START, bool FAST, class Tab, class ClsT>
__device__ __forceinline__ void
walk_line(const ProgramDev &pd, const Tab &tab, const ClsT *cls,
          const uint8_t *p, uint32_t len, uint32_t v0, int32_t &result,
          uint32_t &start, uint32_t &end) {
  Walk<STYLE, WANT_START> wk;
  wk.begin(pd);
  wk.v = v0;

  const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
  const uint32_t skip = static_cast<uint32_t>(addr & 15);
  const uint4 *blk = reinterpret_cast<const uint4 *>(addr - skip);
  const uint32_t total = skip + len;  // bytes from blk[0] to the line's end
  const uint32_t nblk = (total + 15) >> 4;

  if (len != 0) {
    uint4 cur = __ldg(blk);
    for (uint32_t k = 0; k < nblk; ++k) {
      uint4 nxt = cur;
      if (k + 1 < nblk)
        nxt = __ldg(blk + k + 1);
      const uint32_t pos0 = k * 16 - skip;  // line position of byte 0 (mod 2^32)
      const uint32_t lo = (k == 0) ? skip : 0;
      const uint32_t hi = min(16u, total - k * 16);
      if ((lo == 0) && (hi == 16)) {
        if (FAST) {
          // Walk the 16 bytes without looking at what each state means; only
          // remember whether any of them would have needed a decision.  If
          // none did (the common case) the chunk is done.  Otherwise rewind
          // and take the careful steps below.
          Walk<STYLE, WANT_START> sv = wk;
          bool any = false;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            uint32_t b = byte_of(word_of(cur, j >> 2), j & 3);
            uint32_t nv = tab.next(wk.v, cls[b]);
            if (WANT_START) {
              if ((wk.v == pd.initV) && (nv != pd.initV))
                wk.mstart = pos0 + j;
            }
            wk.v = nv;
            any |= (nv >= wk.limit);
          }
          if (!any) {
            cur = nxt;
            continue;
          }
          wk = sv;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t b = byte_of(word_of(cur, j >> 2), j & 3);
          if (wk.step(pd, tab, cls[b], pos0 + j))
            goto done;
        }
      }
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if ((j >= lo) && (j < hi)) {
            uint32_t b = byte_of(word_of(cur, j >> 2), j & 3);
            if (wk.step(pd, tab, cls[b], pos0 + j))
              goto done;
          }
        }
      }
      cur = nxt;
    }
  }
done:
  wk.finish(pd, len, result, start, end);
}

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT, bool FAST>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_anchored_kernel(const ProgramDev pd, const BatchArgs args) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);

  // persistent CTAs: tile t covers lines [t*blockDim, (t+1)*blockDim)
  const uint64_t nLines = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint64_t tiles = (nLines + blockDim.x - 1) / blockDim.x;
  for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const uint64_t line = t * blockDim.x + threadIdx.x;
    if (line >= nLines)
      continue;
    const uint8_t *p = args.text + args.starts[line];
    uint32_t len = args.lens[line];
    uint32_t v0 = pd.initV;
    int32_t result = 0;
    uint32_t start = 0, end = 0;
    bool viable = true;
    if (args.checkLeader) {
      /* check() with the leader optimisation (Matcher.h:369-376): the
         required prefix is compared first and the walk starts after it, from
         the state it leads to.  Kept because it is observable: an accepting
         state reached exactly at the end of the leader is not noticed until
         the next byte has been consumed. */
      for (uint32_t ii = 0; ii < pd.leaderLen; ++ii)
        if ((ii >= len) || (pd.leader[ii] != pd.equiv[__ldg(p + ii)])) {
          viable = false;
          break;
        }
      p += pd.leaderLen;
      len -= pd.leaderLen;
      v0 = pd.leaderV;
    }
    if (!viable) {
      // nothing: result 0
    }
    else if (SMEM) {
      SmemTable<K> tab{smem + kClsBytes};
      walk_line<STYLE, WANT_START, FAST>(pd, tab, cls, p, len, v0, result, start, end);
    }
    else {
      GlobalTable tab{static_cast<const unsigned char *>(pd.table)};
      walk_line<STYLE, WANT_START, FAST>(pd, tab, cls, p, len, v0, result, start, end);
    }
    args.results[line] = result;
    if (WANT_START)
      args.matchStarts[line] = start;
    if (args.matchEnds)
      args.matchEnds[line] = end;
  }
}

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT>
cudaError_t launch_one(const ProgramDev &pd, const BatchArgs &args,
                       const LaunchGeometry &geo, const BatchTuning &tune,
                       cudaStream_t stream) {
  auto kern = tune.simpleWalk
      ? batch_anchored_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>
      : batch_anchored_kernel<STYLE, WANT_START, SMEM, K, ClsT, true>;
  size_t smem = kClsBytes + (SMEM ? pd.tableBytes : 0);
  cudaError_t err = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess)
    return err;
  int threads = tune.threads ? tune.threads : kMaxThreads;
  int perSm = tune.ctasPerSm ? tune.ctasPerSm : 1;
  uint64_t tiles = (args.nLines + threads - 1) / threads;
  uint64_t grid = static_cast<uint64_t>(geo.smCount) * perSm;
  if (grid > tiles)
    grid = tiles;
  if (grid == 0)
    return cudaSuccess;
  kern<<<static_cast<unsigned>(grid), threads, smem, stream>>>(pd, args);
  return cudaGetLastError();
}

template <int STYLE, bool WANT_START>
cudaError_t dispatch_table(const ProgramDev &pd, bool smemTable,
                           const BatchArgs &args, const LaunchGeometry &geo,
                           const BatchTuning &tune, cudaStream_t stream) {
  const bool cls8 = pd.clsMaxOffset < 256;
  if (smemTable) {
    if (pd.unitShift == 0)
      return cls8 ? launch_one<STYLE, WANT_START, true, 1, uint8_t>(pd, args, geo, tune, stream)
                  : launch_one<STYLE, WANT_START, true, 1, uint16_t>(pd, args, geo, tune, stream);
    if (pd.unitShift == 1)
      return cls8 ? launch_one<STYLE, WANT_START, true, 2, uint8_t>(pd, args, geo, tune, stream)
                  : launch_one<STYLE, WANT_START, true, 2, uint16_t>(pd, args, geo, tune, stream);
    if (pd.unitShift == 8)  // byte-indexed rows
      return launch_one<STYLE, WANT_START, true, 9, uint16_t>(pd, args, geo, tune, stream);
    return cudaErrorInvalidValue;
  }
  return cls8 ? launch_one<STYLE, WANT_START, false, 0, uint8_t>(pd, args, geo, tune, stream)
              : launch_one<STYLE, WANT_START, false, 0, uint16_t>(pd, args, geo, tune, stream);
}

template <bool WANT_START>
cudaError_t dispatch_style(const ProgramDev &pd, bool smemTable,
                           const BatchArgs &args, int style,
                           const LaunchGeometry &geo, const BatchTuning &tune,
                           cudaStream_t stream) {
  switch (style) {
  case STY_INSTANT: return dispatch_table<STY_INSTANT, WANT_START>(pd, smemTable, args, geo, tune, stream);
  case STY_FIRST:   return dispatch_table<STY_FIRST, WANT_START>(pd, smemTable, args, geo, tune, stream);
  case STY_TANGENT: return dispatch_table<STY_TANGENT, WANT_START>(pd, smemTable, args, geo, tune, stream);
  case STY_LAST:    return dispatch_table<STY_LAST, WANT_START>(pd, smemTable, args, geo, tune, stream);
  case STY_FULL:    return dispatch_table<STY_FULL, WANT_START>(pd, smemTable, args, geo, tune, stream);
  default:          return cudaErrorInvalidValue;
  }
}


// ---- sliding verbs: scan / search -----------------------------------------

/* scanCore / searchCore (Matcher.h:498-641): try the anchored walk at
   successive start positions until one matches.  One lane per line, bytes
   read one at a time through the read-only path: this is the reference's
   O(len^2) algorithm as is, kept simple because multi-pattern users are told
   to use a loose-start DFA with match() instead (doc/Usage.md).

   SCAN reproduces a quirk of the reference when the program has a leader:
   compareThrough() advances the caller's cursor (Matcher.h:349), so after a
   leader comparison - failed or not - the bytes it consumed are never tried
   as start positions. */
template <int STYLE, bool SEARCH, class Tab, class ClsT>
__device__ __forceinline__ void
slide_line(const ProgramDev &pd, const Tab &tab, const ClsT *cls,
           const uint8_t *p, uint32_t len, int32_t &result, uint32_t &start,
           uint32_t &end) {
  Walk<STYLE, SEARCH> wk;
  wk.begin(pd);
  wk.finish(pd, 0, result, start, end);  // empty text: the initial state's result
  start = 0;
  end = 0;
  if (len == 0)
    return;
  result = 0;
  for (uint32_t pos = 0; pos < len; ++pos) {
    wk.begin(pd);
    if (!SEARCH && (pd.leaderLen > 0)) {
      bool ok = true;
      uint32_t ii = 0;
      for (; ii < pd.leaderLen; ++ii) {
        if ((pos + ii >= len) || (pd.leader[ii] != pd.equiv[__ldg(p + pos + ii)])) {
          ok = false;
          break;
        }
      }
      pos += ii;
      if (!ok)
        continue;
      wk.v = pd.leaderV;
    }
    wk.mstart = pos;
    wk.mend = pos;
    for (uint32_t idx = pos; idx < len; ++idx)
      if (wk.step(pd, tab, cls[__ldg(p + idx)], idx))
        break;
    int32_t r;
    uint32_t s, e;
    wk.finish(pd, len, r, s, e);
    if (r > 0) {
      result = r;
      start = s;
      end = e;
      return;
    }
  }
}

template <int STYLE, bool SEARCH, bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_sliding_kernel(const ProgramDev pd, const BatchArgs args) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const uint64_t nLines = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint64_t tiles = (nLines + blockDim.x - 1) / blockDim.x;
  for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const uint64_t line = t * blockDim.x + threadIdx.x;
    if (line >= nLines)
      continue;
    const uint8_t *p = args.text + args.starts[line];
    const uint32_t len = args.lens[line];
    int32_t result;
    uint32_t start, end;
    if (SMEM) {
      SmemTable<K> tab{smem + kClsBytes};
      slide_line<STYLE, SEARCH>(pd, tab, cls, p, len, result, start, end);
    }
    else {
      GlobalTable tab{static_cast<const unsigned char *>(pd.table)};
      slide_line<STYLE, SEARCH>(pd, tab, cls, p, len, result, start, end);
    }
    args.results[line] = result;
    if (args.matchStarts)
      args.matchStarts[line] = start;
    if (args.matchEnds)
      args.matchEnds[line] = end;
  }
}

/* The same verbs with a WARP per line: lane l tries start position base + l,
   32 positions at a time; the lowest lane whose walk matched wins (the
   attempts are independent, so this is the reference's "first start position
   that matches").  Most walks die within a byte or two, so a round is a few
   steps with all lanes busy in the first, and the byte loads of a step are
   coalesced (adjacent lanes read adjacent bytes).  Not used for scan() on a
   program with a leader: there the reference skips start positions
   sequentially (see slide_line). */
template <int STYLE, bool SEARCH, bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_sliding_warp_kernel(const ProgramDev pd, const BatchArgs args,
                          unsigned long long *lineCounter) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes
                     : static_cast<const unsigned char *>(pd.table)};
  const uint32_t lane = threadIdx.x & 31;
  // per warp: up to 63 start positions that wait to be walked, and whose line they are
  uint32_t *surv = reinterpret_cast<uint32_t *>(smem + kClsBytes + (SMEM ? pd.tableBytes : 0)) + (threadIdx.x >> 5) * 80;
  uint8_t *survLine = reinterpret_cast<uint8_t *>(surv + 64);
  const uint64_t nLines = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  constexpr uint32_t kBatch = 8;  // lines claimed at a time
  for (;;) {
    unsigned long long first = 0;
    if (lane == 0)
      first = atomicAdd(lineCounter, static_cast<unsigned long long>(kBatch));
    first = __shfl_sync(0xffffffffu, first, 0);
    if (first >= nLines)
      break;
    const uint64_t last = min(nLines, static_cast<uint64_t>(first) + kBatch);
    // the batch's index entries are loaded together (lane j holds line j's) and
    // the head of every line is pulled towards the SM before the first is walked
    uint64_t myStart = 0;
    uint32_t myLen = 0;
    if (first + lane < last) {
      myStart = __ldg(args.starts + first + lane);
      myLen = __ldg(args.lens + first + lane);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(args.text + myStart));
      if (myLen > 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(args.text + myStart + 128));
    }
    /* Filter, then walk.  Nearly every start position dies within two bytes
       (pd.startFilter: one bit per byte pair); a round of 32 positions that
       walks them all lasts as long as its longest walk with a third of the
       lanes busy.  So a round only LOOKS at its positions; the ones that may
       match - of all the lines of the batch, in order - are collected, and
       whenever 32 of them wait (and at the batch's end) they are walked
       together, one per lane.  The lowest lane of a line that matched is the
       first position of that line that does; a line that has its match takes
       no further part.  (Candidates of several lines in one round: the long
       walk of a true match then has thirty other walks beside it, not five.) */
    const uint32_t nb = static_cast<uint32_t>(last - first);
    int32_t myRes = 0;        // lane j: the outcome of line first + j
    uint32_t myS = 0, myE = 0;
    uint32_t foundMask = 0;   // lines that have their match (warp-uniform)
    uint32_t nSurv = 0;
    auto walkGroup = [&](uint32_t count) {
      int32_t r = 0;
      uint32_t sPos = 0, e = 0, j = 0;
      if (lane < count) {
        j = survLine[lane];
      }
      const uint64_t st = __shfl_sync(0xffffffffu, myStart, static_cast<int>(j));
      const uint32_t len = __shfl_sync(0xffffffffu, myLen, static_cast<int>(j));
      if ((lane < count) && !((foundMask >> j) & 1u)) {
        const uint8_t *p = args.text + st;
        const uint32_t pos = surv[lane];
        Walk<STYLE, SEARCH> wk;
        wk.begin(pd);
        wk.mstart = pos;
        wk.mend = pos;
        for (uint32_t idx = pos; idx < len; ++idx)
          if (wk.step(pd, tab, cls[__ldg(p + idx)], idx))
            break;
        wk.finish(pd, len, r, sPos, e);
      }
      for (uint32_t L = 0; L < nb; ++L) {
        const unsigned hits = __ballot_sync(0xffffffffu, (r > 0) && (j == L) && (lane < count));
        if (hits && !((foundMask >> L) & 1u)) {
          const int w = __ffs(hits) - 1;
          const int32_t rW = __shfl_sync(0xffffffffu, r, w);
          const uint32_t sW = __shfl_sync(0xffffffffu, sPos, w);
          const uint32_t eW = __shfl_sync(0xffffffffu, e, w);
          if (lane == L) {
            myRes = rW;
            myS = sW;
            myE = eW;
          }
          foundMask |= 1u << L;
        }
      }
    };
    for (uint32_t j = 0; j < nb; ++j) {
      const uint8_t *p = args.text + __shfl_sync(0xffffffffu, myStart, static_cast<int>(j));
      const uint32_t len = __shfl_sync(0xffffffffu, myLen, static_cast<int>(j));
      if (len == 0) {
        // empty text: the initial state's result, positions 0
        Walk<STYLE, SEARCH> wk;
        wk.begin(pd);
        int32_t r0 = 0;
        uint32_t s0 = 0, e0 = 0;
        wk.finish(pd, 0, r0, s0, e0);
        if (lane == j)
          myRes = r0;
      }
      for (uint32_t base = 0; (base < len) && !((foundMask >> j) & 1u); base += 32) {
        const uint32_t pos = base + lane;
        bool alive = false;
        if (pos < len) {
          const uint32_t b0 = __ldg(p + pos);
          const uint32_t bit = (pos + 1 < len) ? ((b0 << 8) | __ldg(p + pos + 1)) : (65536u + b0);
          alive = (__ldg(pd.startFilter + (bit >> 5)) >> (bit & 31)) & 1u;
        }
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (alive) {
          const uint32_t slot = nSurv + __popc(m & ((1u << lane) - 1));
          surv[slot] = pos;
          survLine[slot] = static_cast<uint8_t>(j);
        }
        nSurv += __popc(m);
        __syncwarp();
        if (nSurv >= 32) {
          walkGroup(32);
          __syncwarp();
          // the rest (at most 31) moves to the front
          const uint32_t left = nSurv - 32;
          const uint32_t movedPos = (lane < left) ? surv[32 + lane] : 0;
          const uint8_t movedLine = (lane < left) ? survLine[32 + lane] : 0;
          __syncwarp();
          if (lane < left) {
            surv[lane] = movedPos;
            survLine[lane] = movedLine;
          }
          nSurv = left;
          __syncwarp();
        }
      }
    }
    if (nSurv)
      walkGroup(nSurv);
    __syncwarp();
    if (lane < nb) {
      const uint64_t line = first + lane;
      args.results[line] = myRes;
      if (args.matchStarts)
        args.matchStarts[line] = myS;
      if (args.matchEnds)
        args.matchEnds[line] = myE;
    }
  }
}

template <int STYLE, bool SEARCH, bool SMEM, int K, class ClsT>
cudaError_t launch_slide(const ProgramDev &pd, const BatchArgs &args,
                         const LaunchGeometry &geo, cudaStream_t stream,
                         unsigned long long *counter, bool simple) {
  auto kern = batch_sliding_kernel<STYLE, SEARCH, SMEM, K, ClsT>;
  size_t smem = kClsBytes + (SMEM ? pd.tableBytes : 0);
  cudaError_t err = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess)
    return err;
  if (counter && !simple && (SEARCH || (pd.leaderLen == 0))) {
    auto wkern = batch_sliding_warp_kernel<STYLE, SEARCH, SMEM, K, ClsT>;
    const size_t wsmem = smem + (kMaxThreads / 32) * 320;  // + the warps' lists of waiting positions
    err = cudaFuncSetAttribute(wkern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(wsmem));
    if (err != cudaSuccess)
      return err;
    if (args.nLines == 0)
      return cudaSuccess;
    err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess)
      return err;
    // one CTA per SM when the table takes most of the shared memory
    const uint64_t perSmW = (SMEM && (smem > (100u << 10))) ? 1 : 2;
    const int wthreads = (perSmW == 1) ? kMaxThreads : 512;
    const uint64_t warpsNeeded = (args.nLines + 7) / 8;
    const uint64_t wgrid = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * perSmW,
                                              (warpsNeeded + wthreads / 32 - 1) / (wthreads / 32));
    wkern<<<static_cast<unsigned>(wgrid), wthreads, wsmem, stream>>>(pd, args, counter);
    return cudaGetLastError();
  }
  const int threads = 256;  // lines differ wildly in cost: smaller tiles
  uint64_t tiles = (args.nLines + threads - 1) / threads;
  uint64_t perSm = SMEM ? std::max<uint64_t>(1, std::min<uint64_t>(4, (200u << 10) / smem)) : 4;
  uint64_t grid = std::min<uint64_t>(tiles, static_cast<uint64_t>(geo.smCount) * perSm);
  if (grid == 0)
    return cudaSuccess;
  kern<<<static_cast<unsigned>(grid), threads, smem, stream>>>(pd, args);
  return cudaGetLastError();
}

template <int STYLE, bool SEARCH>
cudaError_t slide_table(const ProgramDev &pd, bool smemTable,
                        const BatchArgs &args, const LaunchGeometry &geo,
                        cudaStream_t stream, unsigned long long *counter, bool simple) {
  const bool cls8 = pd.clsMaxOffset < 256;
  if (smemTable) {
    if (pd.unitShift == 0)
      return cls8 ? launch_slide<STYLE, SEARCH, true, 1, uint8_t>(pd, args, geo, stream, counter, simple)
                  : launch_slide<STYLE, SEARCH, true, 1, uint16_t>(pd, args, geo, stream, counter, simple);
    if (pd.unitShift == 1)
      return cls8 ? launch_slide<STYLE, SEARCH, true, 2, uint8_t>(pd, args, geo, stream, counter, simple)
                  : launch_slide<STYLE, SEARCH, true, 2, uint16_t>(pd, args, geo, stream, counter, simple);
    if (pd.unitShift == 8)
      return launch_slide<STYLE, SEARCH, true, 9, uint16_t>(pd, args, geo, stream, counter, simple);
    return cudaErrorInvalidValue;
  }
  return cls8 ? launch_slide<STYLE, SEARCH, false, 0, uint8_t>(pd, args, geo, stream, counter, simple)
              : launch_slide<STYLE, SEARCH, false, 0, uint16_t>(pd, args, geo, stream, counter, simple);
}

template <bool SEARCH>
cudaError_t slide_style(const ProgramDev &pd, bool smemTable,
                        const BatchArgs &args, int style,
                        const LaunchGeometry &geo, cudaStream_t stream,
                        unsigned long long *counter, bool simple) {
  switch (style) {
  case STY_INSTANT: return slide_table<STY_INSTANT, SEARCH>(pd, smemTable, args, geo, stream, counter, simple);
  case STY_FIRST:   return slide_table<STY_FIRST, SEARCH>(pd, smemTable, args, geo, stream, counter, simple);
  case STY_TANGENT: return slide_table<STY_TANGENT, SEARCH>(pd, smemTable, args, geo, stream, counter, simple);
  case STY_LAST:    return slide_table<STY_LAST, SEARCH>(pd, smemTable, args, geo, stream, counter, simple);
  case STY_FULL:    return slide_table<STY_FULL, SEARCH>(pd, smemTable, args, geo, stream, counter, simple);
  default:          return cudaErrorInvalidValue;
  }
}

} // anonymous

cudaError_t launch_batch_anchored(const ProgramDev &pd, bool smemTable,
                                  const BatchArgs &args, int style,
                                  const LaunchGeometry &geo,
                                  const BatchTuning &tune, cudaStream_t stream,
                                  int *launches) {
  cudaError_t err = args.matchStarts
      ? dispatch_style<true>(pd, smemTable, args, style, geo, tune, stream)
      : dispatch_style<false>(pd, smemTable, args, style, geo, tune, stream);
  if ((err == cudaSuccess) && launches && args.nLines)
    *launches += 1;
  return err;
}

cudaError_t launch_batch_sliding(const ProgramDev &pd, bool smemTable,
                                 const BatchArgs &args, int style,
                                 bool wantPositions,
                                 const LaunchGeometry &geo,
                                 cudaStream_t stream, int *launches,
                                 unsigned long long *counter, bool simple) {
  cudaError_t err = wantPositions
      ? slide_style<true>(pd, smemTable, args, style, geo, stream, counter, simple)
      : slide_style<false>(pd, smemTable, args, style, geo, stream, counter, simple);
  if ((err == cudaSuccess) && launches && args.nLines)
    *launches += 1;
  return err;
}

} // namespace red_b200
