/* kernels_multi.cu - per line, ALL matches instead of one:

     which = 0  matchAll()       one anchored walk that reports every run of
                                 positions with the same positive result
                                 (reference: matchAllCore, Matcher.h:711-767)
     which = 1  Red::collect()   repeated search<styLast> from just after the
                                 previous match: all non-overlapping matches in
                                 order (reference: lib/Red.cpp:108-122)

   One lane per line; spans[line * maxPerLine + k] holds the k-th outcome,
   counts[line] the number found (it can exceed maxPerLine; only the first
   maxPerLine are stored - the caller sees that from the count).  Positions
   are relative to the line, as Outcome's are to the text. */

#include "device_program.h"
#include "walk.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace red_b200 {

namespace {

using namespace dev;

template <class Tab, class ClsT>
__device__ __forceinline__ uint32_t match_all_line(const ProgramDev &pd, const Tab &tab,
                                                   const ClsT *cls, const uint8_t *p,
                                                   uint32_t len, uint32_t cap,
                                                   MultiSpan *out) {
  // the leader only gates (Matcher.h:724-729)
  for (uint32_t ii = 0; ii < pd.leaderLen; ++ii)
    if ((ii >= len) || (pd.leader[ii] != pd.equiv[__ldg(p + ii)]))
      return 0;
  uint32_t v = pd.initV, cnt = 0, mstart = 0;
  int32_t prev = 0;
  for (uint32_t idx = 0; idx < len; ++idx) {
    const uint32_t nv = tab.next(v, cls[__ldg(p + idx)]);
    if ((v == pd.initV) && (nv != pd.initV))
      mstart = idx;
    v = nv;
    if (v >= pd.acceptV) {
      if (v >= pd.pureV)
        break;
      const int32_t r = pd.results[state_index(pd, v)];
      if (r == prev) {
        if (cnt - 1 < cap)
          out[cnt - 1].end = idx + 1;
      }
      else {
        prev = r;
        if (cnt < cap) {
          out[cnt].result = r;
          out[cnt].start = mstart;
          out[cnt].end = idx + 1;
        }
        ++cnt;
      }
    }
    else
      prev = 0;
  }
  return cnt;
}

template <class Tab, class ClsT>
__device__ __forceinline__ uint32_t collect_line(const ProgramDev &pd, const Tab &tab,
                                                 const ClsT *cls, const uint8_t *p,
                                                 uint32_t len, uint32_t cap,
                                                 MultiSpan *out) {
  uint32_t cnt = 0, off = 0;
  while (off < len) {
    // search<styLast, false> on p[off..len)
    int32_t result = 0;
    uint32_t start = 0, end = 0;
    for (uint32_t pos = off; (pos < len) && (result == 0); ++pos) {
      Walk<STY_LAST, true> wk;
      wk.begin(pd);
      wk.mstart = pos;
      wk.mend = pos;
      for (uint32_t idx = pos; idx < len; ++idx)
        if (wk.step(pd, tab, cls[__ldg(p + idx)], idx))
          break;
      wk.finish(pd, len, result, start, end);
    }
    if (result <= 0)
      break;
    if (cnt < cap) {
      out[cnt].result = result;
      out[cnt].start = start;
      out[cnt].end = end;
    }
    ++cnt;
    off = end;
  }
  return cnt;
}

template <bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(256)
batch_multi_kernel(const ProgramDev pd, const uint8_t *text, const uint64_t *starts,
                   const uint32_t *lens, uint64_t nLines, int which,
                   uint32_t maxPerLine, uint32_t *counts, MultiSpan *spans) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes
                     : static_cast<const unsigned char *>(pd.table)};
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t line = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       line < nLines; line += stride) {
    const uint8_t *p = text + starts[line];
    const uint32_t len = lens[line];
    MultiSpan *out = spans + line * maxPerLine;
    counts[line] = (which == 0)
        ? match_all_line(pd, tab, cls, p, len, maxPerLine, out)
        : collect_line(pd, tab, cls, p, len, maxPerLine, out);
  }
}

/* replace() for one line (reference: replaceCore, Matcher.h:643-707): every
   position in order; a match is substituted and skipped, anything else is
   copied through; after maxRepl substitutions the rest is copied.  The leader
   test of the reference (lookingAt, :663) only avoids walks that would fail on
   their own, so it is not repeated here.  WRITE = false only measures. */
template <bool WRITE, class Tab, class ClsT>
__device__ __forceinline__ uint32_t replace_line(const ProgramDev &pd, const Tab &tab,
                                                 const ClsT *cls, const uint8_t *p,
                                                 uint32_t len, int style, uint64_t maxRepl,
                                                 const uint8_t *repl, uint32_t replLen,
                                                 uint8_t *out, uint32_t &outLen) {
  uint32_t cnt = 0, o = 0, pos = 0;
  while (pos < len) {
    if (cnt >= maxRepl) {
      if (WRITE)
        for (uint32_t ii = pos; ii < len; ++ii)
          out[o + ii - pos] = __ldg(p + ii);
      o += len - pos;
      break;
    }
    bool have = false;
    uint32_t found = 0;
    uint32_t v = pd.initV;
    int32_t prev = 0;
    for (uint32_t idx = pos; idx < len; ++idx) {
      v = tab.next(v, cls[__ldg(p + idx)]);
      if ((v >= pd.acceptV) && (v < pd.pureV)) {
        if (style == STY_FIRST) {
          const int32_t r = pd.results[state_index(pd, v)];
          if (prev && (r != prev))
            break;
          prev = r;
        }
        have = true;
        found = idx;
        if (style == STY_INSTANT)
          break;
        if (v >= pd.absV) {
          // can never be left: every later position accepts with this result
          found = len - 1;
          break;
        }
      }
      else {
        if (style == STY_FULL)
          have = false;
        if ((((style == STY_FIRST) || (style == STY_TANGENT)) && have) || (v >= pd.pureV))
          break;
      }
    }
    if (have) {
      if (WRITE)
        for (uint32_t ii = 0; ii < replLen; ++ii)
          out[o + ii] = __ldg(repl + ii);
      o += replLen;
      pos = found + 1;
      ++cnt;
    }
    else {
      if (WRITE)
        out[o] = __ldg(p + pos);
      ++o;
      ++pos;
    }
  }
  outLen = o;
  return cnt;
}

/* pass 0 (WRITE = false): counts[line], outLens[line].
   pass 1 (WRITE = true): the rewritten line at outText + outStarts[line],
   followed by the delimiter if one is asked for. */
template <bool SMEM, int K, class ClsT, bool WRITE>
__global__ void __launch_bounds__(256)
batch_replace_kernel(const ProgramDev pd, const uint8_t *text, const uint64_t *starts,
                     const uint32_t *lens, uint64_t nLines, int style, uint64_t maxRepl,
                     const uint8_t *repl, uint32_t replLen, int delim, uint32_t *counts,
                     uint32_t *outLens, const uint64_t *outStarts, uint8_t *outText) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes
                     : static_cast<const unsigned char *>(pd.table)};
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t line = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       line < nLines; line += stride) {
    const uint8_t *p = text + starts[line];
    const uint32_t len = lens[line];
    uint32_t outLen = 0;
    uint8_t *out = WRITE ? outText + outStarts[line] : nullptr;
    const uint32_t cnt = replace_line<WRITE>(pd, tab, cls, p, len, style, maxRepl, repl,
                                             replLen, out, outLen);
    if (WRITE) {
      if (delim >= 0)
        out[outLen] = static_cast<uint8_t>(delim);
    }
    else {
      counts[line] = cnt;
      outLens[line] = outLen;
    }
  }
}

/* replace() with a WARP per line: the 32 lanes try 32 consecutive start
   positions at once; the lowest lane whose walk found a match wins (the
   attempts are independent, so this is the reference's "first position that
   matches"), the text in front of it is copied through and the search goes on
   behind the match.  Most walks die within a byte or two and the byte loads of
   a step are coalesced; copying is done by the whole warp.  Same two passes
   as the lane-per-line kernel (measure, then write). */
constexpr uint32_t kReplaceBuf = 1024;   // bytes of line text per warp in shared memory
constexpr uint32_t kReplaceList = 256;   // start positions per line that may match (more: the plain loop)
constexpr uint32_t kReplaceWarpBytes = kReplaceBuf + 4 * kReplaceList;  // text, positions (u16), ends (u16)

template <bool SMEM, int K, class ClsT, bool WRITE>
__global__ void __launch_bounds__(512)
batch_replace_warp_kernel(const ProgramDev pd, const uint8_t *text, const uint64_t *starts,
                          const uint32_t *lens, uint64_t nLines, int style, uint64_t maxRepl,
                          const uint8_t *repl, uint32_t replLen, int delim, uint32_t *counts,
                          uint32_t *outLens, const uint64_t *outStarts, uint8_t *outText) {
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
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nWarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  /* A line is read many times over (every start position is a walk of its
     own), byte by byte, each byte a load the next step waits for: the warp
     copies the line into shared memory first - all of its loads in flight at
     once - and walks it from there; the next line's text is asked into L2
     meanwhile.  A line longer than the buffer is walked from global memory. */
  uint8_t *wbuf = smem + kClsBytes + (SMEM ? pd.tableBytes : 0) + (threadIdx.x >> 5) * kReplaceWarpBytes;
  uint16_t *sPos = reinterpret_cast<uint16_t *>(wbuf + kReplaceBuf);
  uint16_t *sEnd = sPos + kReplaceList;
  for (uint64_t line = warp; line < nLines; line += nWarps) {
    const uint8_t *p = text + __ldg(starts + line);
    const uint32_t len = __ldg(lens + line);
    if ((line + nWarps < nLines) && (lane < 2))
      asm volatile("prefetch.global.L2 [%0];" ::"l"(text + __ldg(starts + line + nWarps) + 128 * lane));
    if (len <= kReplaceBuf) {
      __syncwarp();
      for (uint32_t ii = lane; ii < len; ii += 32)
        wbuf[ii] = __ldg(p + ii);
      __syncwarp();
      p = wbuf;
    }
    uint8_t *out = WRITE ? outText + __ldg(outStarts + line) : nullptr;
    uint32_t cnt = 0, o = 0, pos = 0;
    auto copyThrough = [&](uint32_t from, uint32_t to) {  // text[from, to) -> out + o
      if (WRITE)
        for (uint32_t ii = from + lane; ii < to; ii += 32)
          out[o + ii - from] = p[ii];
      o += to - from;
    };
    // the anchored walk from `at` (replaceCore's inner loop): is there a match, and its last byte
    auto walkFrom = [&](uint32_t at, uint32_t &found) -> bool {
      bool have = false;
      uint32_t v = pd.initV;
      int32_t prev = 0;
      for (uint32_t idx = at; idx < len; ++idx) {
        v = tab.next(v, cls[p[idx]]);
        if ((v >= pd.acceptV) && (v < pd.pureV)) {
          if (style == STY_FIRST) {
            const int32_t r = pd.results[state_index(pd, v)];
            if (prev && (r != prev))
              break;
            prev = r;
          }
          have = true;
          found = idx;
          if (style == STY_INSTANT)
            break;
          if (v >= pd.absV) {
            found = len - 1;
            break;
          }
        }
        else {
          if (style == STY_FULL)
            have = false;
          if ((((style == STY_FIRST) || (style == STY_TANGENT)) && have) || (v >= pd.pureV))
            break;
        }
      }
      return have;
    };
    // (pd.startFilter: the bytes at a position say whether a walk from it can match at all)
    auto mayMatch = [&](uint32_t at) -> bool {
      const uint32_t b0 = p[at];
      const uint32_t bit = (at + 1 < len) ? ((b0 << 8) | p[at + 1]) : (65536u + b0);
      return (__ldg(pd.startFilter + (bit >> 5)) >> (bit & 31)) & 1u;
    };
    auto substitute = [&](uint32_t at, uint32_t end) {
      copyThrough(pos, at);
      if (WRITE)
        for (uint32_t ii = lane; ii < replLen; ii += 32)
          out[o + ii] = __ldg(repl + ii);
      o += replLen;
      pos = end;
      ++cnt;
    };

    /* A staged line in three steps.  (1) Every position is LOOKED at, the few
       that may match go on a list, in order.  (2) They are walked 32 at a
       time, one per lane - the walks are independent of each other and of what
       gets substituted: each starts in the initial state and reads the line as
       it is.  (3) The substitutions are read off in order: the first listed
       position at or behind the cursor that has a match, then the first one at
       or behind that match's end, and so on.  The same result as trying every
       position in turn, without a round of 32 walks per 32 bytes of text. */
    bool listed = false;
    if ((p == wbuf) && (maxRepl > 0)) {
      uint32_t nSurv = 0;
      for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t at = base + lane;
        const bool alive = (at < len) && mayMatch(at);
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        const uint32_t slot = nSurv + __popc(m & ((1u << lane) - 1));
        if (alive && (slot < kReplaceList))
          sPos[slot] = static_cast<uint16_t>(at);
        nSurv += __popc(m);
      }
      __syncwarp();
      if (nSurv <= kReplaceList) {
        listed = true;
        for (uint32_t base = 0; base < nSurv; base += 32) {
          const uint32_t i = base + lane;
          if (i < nSurv) {
            uint32_t found = 0;
            const bool have = walkFrom(sPos[i], found);
            sEnd[i] = have ? static_cast<uint16_t>(found + 1) : 0;
          }
        }
        __syncwarp();
        for (uint32_t base = 0; (base < nSurv) && (cnt < maxRepl);) {
          const uint32_t i = base + lane;
          const bool cand = (i < nSurv) && (sEnd[i] != 0) && (sPos[i] >= pos);
          const unsigned m = __ballot_sync(0xffffffffu, cand);
          if (!m) {
            base += 32;
            continue;
          }
          const uint32_t w = base + (__ffs(m) - 1);
          substitute(sPos[w], sEnd[w]);
          base = w + 1;
        }
        copyThrough(pos, len);
        pos = len;
      }
    }
    while (!listed && (pos < len)) {
      if (cnt >= maxRepl) {
        copyThrough(pos, len);
        break;
      }
      // lane l: is there a match starting at pos + l, and where does it end
      const uint32_t at = pos + lane;
      uint32_t found = 0;
      const bool have = (at < len) && mayMatch(at) && walkFrom(at, found);
      const unsigned hits = __ballot_sync(0xffffffffu, have);
      if (hits) {
        const int w = __ffs(hits) - 1;
        substitute(pos + w, __shfl_sync(0xffffffffu, found, w) + 1);
      }
      else {
        const uint32_t to = min(len, pos + 32);
        copyThrough(pos, to);
        pos = to;
      }
    }
    if (lane == 0) {
      if (WRITE) {
        if (delim >= 0)
          out[o] = static_cast<uint8_t>(delim);
      }
      else {
        counts[line] = cnt;
        outLens[line] = o;
      }
    }
  }
}

/* StatefulMatcher::advance over a segment per line (Matcher.h:770-794): the
   state comes from and goes back to states[line]; kStateFresh starts a new
   matcher.  results[line] is result() after the segment (for an empty segment
   the result of the state itself); hitResults/hitEnds describe the last
   accepting position inside the segment (0/0: none). */
constexpr uint32_t kStateFresh = 0xffffffffu;

template <bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(256)
batch_advance_kernel(const ProgramDev pd, const uint8_t *text, const uint64_t *starts,
                     const uint32_t *lens, uint64_t nLines, uint32_t *states,
                     int32_t *results, int32_t *hitResults, uint32_t *hitEnds) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes
                     : static_cast<const unsigned char *>(pd.table)};
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t line = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       line < nLines; line += stride) {
    const uint8_t *p = text + starts[line];
    const uint32_t len = lens[line];
    uint32_t v = states[line];
    // a state of this program is a multiple of vPerState below stateCount *
    // vPerState; anything else (RED_B200_STATE_FRESH, a stale or foreign
    // value) starts a new matcher instead of indexing the table with it
    if ((v >= pd.stateCount * pd.vPerState) || (state_index(pd, v) * pd.vPerState != v))
      v = pd.initV;
    uint32_t hitV = kNone, hitEnd = 0;
    for (uint32_t idx = 0; idx < len; ++idx) {
      v = tab.next(v, cls[__ldg(p + idx)]);
      if (v >= pd.acceptV) {
        if (v >= pd.absV) {
          // never left again: pure dead end, or accepting for the rest
          if (v < pd.pureV) {
            hitV = v;
            hitEnd = len;
          }
          break;
        }
        hitV = v;
        hitEnd = idx + 1;
      }
    }
    states[line] = v;
    results[line] = ((v >= pd.acceptV) && (v < pd.pureV)) ? pd.results[state_index(pd, v)] : 0;
    if (hitResults)
      hitResults[line] = (hitV != kNone) ? pd.results[state_index(pd, hitV)] : 0;
    if (hitEnds)
      hitEnds[line] = hitEnd;
  }
}

template <bool SMEM, int K, class ClsT>
cudaError_t launch_multi(const ProgramDev &pd, const uint8_t *text, const uint64_t *starts,
                         const uint32_t *lens, uint64_t nLines, int which,
                         uint32_t maxPerLine, uint32_t *counts, MultiSpan *spans,
                         const LaunchGeometry &geo, cudaStream_t stream) {
  auto kern = batch_multi_kernel<SMEM, K, ClsT>;
  size_t smem = kClsBytes + (SMEM ? pd.tableBytes : 0);
  cudaError_t err = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess)
    return err;
  uint64_t perSm = SMEM ? std::max<uint64_t>(1, std::min<uint64_t>(4, (200u << 10) / smem)) : 4;
  uint64_t grid = std::min<uint64_t>((nLines + 255) / 256, static_cast<uint64_t>(geo.smCount) * perSm);
  if (grid == 0)
    return cudaSuccess;
  kern<<<static_cast<unsigned>(grid), 256, smem, stream>>>(pd, text, starts, lens, nLines, which,
                                                           maxPerLine, counts, spans);
  return cudaGetLastError();
}

template <bool SMEM, int K, class ClsT, class Kern, class... Args>
cudaError_t launch_line_kernel(Kern kern, const ProgramDev &pd, uint64_t nLines,
                               const LaunchGeometry &geo, cudaStream_t stream, Args... args) {
  size_t smem = kClsBytes + (SMEM ? pd.tableBytes : 0);
  cudaError_t err = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess)
    return err;
  uint64_t perSm = SMEM ? std::max<uint64_t>(1, std::min<uint64_t>(4, (200u << 10) / smem)) : 4;
  uint64_t grid = std::min<uint64_t>((nLines + 255) / 256, static_cast<uint64_t>(geo.smCount) * perSm);
  if (grid == 0)
    return cudaSuccess;
  kern<<<static_cast<unsigned>(grid), 256, smem, stream>>>(pd, args...);
  return cudaGetLastError();
}

// one warp per line, 512 threads per CTA
template <bool SMEM, int K, class ClsT, class Kern, class... Args>
cudaError_t launch_warp_kernel(Kern kern, const ProgramDev &pd, uint64_t nLines,
                               const LaunchGeometry &geo, cudaStream_t stream, Args... args) {
  size_t smem = kClsBytes + (SMEM ? pd.tableBytes : 0) + (512 / 32) * kReplaceWarpBytes;  // + the warps' line buffers and lists
  cudaError_t err = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess)
    return err;
  const uint64_t perSm = (SMEM && (smem > (100u << 10))) ? 1 : 2;
  const uint64_t grid = std::min<uint64_t>((nLines + 15) / 16, static_cast<uint64_t>(geo.smCount) * perSm);
  if (grid == 0)
    return cudaSuccess;
  kern<<<static_cast<unsigned>(grid), 512, smem, stream>>>(pd, args...);
  return cudaGetLastError();
}

} // anonymous

cudaError_t launch_batch_multi(const ProgramDev &pd, bool smemTable, const uint8_t *text,
                               const uint64_t *starts, const uint32_t *lens, uint64_t nLines,
                               int which, uint32_t maxPerLine, uint32_t *counts,
                               MultiSpan *spans, const LaunchGeometry &geo,
                               cudaStream_t stream, int *launches) {
  const bool cls8 = pd.clsMaxOffset < 256;
  cudaError_t err;
#define GO(S, KK, CT) launch_multi<S, KK, CT>(pd, text, starts, lens, nLines, which, maxPerLine, counts, spans, geo, stream)
  if (smemTable) {
    if (pd.unitShift == 0)
      err = cls8 ? GO(true, 1, uint8_t) : GO(true, 1, uint16_t);
    else if (pd.unitShift == 1)
      err = cls8 ? GO(true, 2, uint8_t) : GO(true, 2, uint16_t);
    else if (pd.unitShift == 8)
      err = GO(true, 9, uint16_t);
    else
      err = cudaErrorInvalidValue;
  }
  else
    err = cls8 ? GO(false, 0, uint8_t) : GO(false, 0, uint16_t);
#undef GO
  if ((err == cudaSuccess) && launches && nLines)
    *launches += 1;
  return err;
}

} // namespace red_b200

namespace red_b200 {

cudaError_t launch_batch_replace(const ProgramDev &pd, bool smemTable, bool write,
                                 const uint8_t *text, const uint64_t *starts,
                                 const uint32_t *lens, uint64_t nLines, int style,
                                 uint64_t maxRepl, const uint8_t *repl, uint32_t replLen,
                                 int delim, uint32_t *counts, uint32_t *outLens,
                                 const uint64_t *outStarts, uint8_t *outText,
                                 const LaunchGeometry &geo, cudaStream_t stream, int *launches) {
  const bool cls8 = pd.clsMaxOffset < 256;
  cudaError_t err;
  const bool perLane = getenv("RED_B200_REPLACE_LANE_PER_LINE") != nullptr;  // the first kernel, for A/B
#define GO2(S, KK, CT, W) (perLane \
    ? launch_line_kernel<S, KK, CT>(batch_replace_kernel<S, KK, CT, W>, pd, nLines, geo, stream, \
        text, starts, lens, nLines, style, maxRepl, repl, replLen, delim, counts, outLens, outStarts, outText) \
    : launch_warp_kernel<S, KK, CT>(batch_replace_warp_kernel<S, KK, CT, W>, pd, nLines, geo, stream, \
        text, starts, lens, nLines, style, maxRepl, repl, replLen, delim, counts, outLens, outStarts, outText))
#define GO(S, KK, CT) (write ? GO2(S, KK, CT, true) : GO2(S, KK, CT, false))
  if (smemTable) {
    if (pd.unitShift == 0)
      err = cls8 ? GO(true, 1, uint8_t) : GO(true, 1, uint16_t);
    else if (pd.unitShift == 1)
      err = cls8 ? GO(true, 2, uint8_t) : GO(true, 2, uint16_t);
    else if (pd.unitShift == 8)
      err = GO(true, 9, uint16_t);
    else
      err = cudaErrorInvalidValue;
  }
  else
    err = cls8 ? GO(false, 0, uint8_t) : GO(false, 0, uint16_t);
#undef GO
#undef GO2
  if ((err == cudaSuccess) && launches && nLines)
    *launches += 1;
  return err;
}

cudaError_t launch_batch_advance(const ProgramDev &pd, bool smemTable, const uint8_t *text,
                                 const uint64_t *starts, const uint32_t *lens, uint64_t nLines,
                                 uint32_t *states, int32_t *results, int32_t *hitResults,
                                 uint32_t *hitEnds, const LaunchGeometry &geo,
                                 cudaStream_t stream, int *launches) {
  const bool cls8 = pd.clsMaxOffset < 256;
  cudaError_t err;
#define GO(S, KK, CT) launch_line_kernel<S, KK, CT>(batch_advance_kernel<S, KK, CT>, pd, nLines, geo, stream, \
    text, starts, lens, nLines, states, results, hitResults, hitEnds)
  if (smemTable) {
    if (pd.unitShift == 0)
      err = cls8 ? GO(true, 1, uint8_t) : GO(true, 1, uint16_t);
    else if (pd.unitShift == 1)
      err = cls8 ? GO(true, 2, uint8_t) : GO(true, 2, uint16_t);
    else if (pd.unitShift == 8)
      err = GO(true, 9, uint16_t);
    else
      err = cudaErrorInvalidValue;
  }
  else
    err = cls8 ? GO(false, 0, uint8_t) : GO(false, 0, uint16_t);
#undef GO
  if ((err == cudaSuccess) && launches && nLines)
    *launches += 1;
  return err;
}

} // namespace red_b200
