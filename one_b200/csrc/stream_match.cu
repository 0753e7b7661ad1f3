/* stream_match.cu - match()/check() of ONE long byte stream
   (reference: matchCore over a single text, include/Matcher.h:413-495).

   A DFA walk is a chain of dependent lookups, so one stream has no
   parallelism unless the start state of a piece is known before the pieces
   in front of it are done.  What is done here:

     1. cut the text into chunks (one lane per chunk, ~16 KB each);
     2. SPECULATE the state at the start of chunk i: run the DFA from the
        initial state over a lookback window in front of the chunk.  DFAs of
        the ".*pattern" kind forget their past quickly, so the guess is almost
        always the true state;
     3. walk every chunk from its guess, in parallel, recording the chunk's
        end state and what the styles need: first and last accepting
        position (with the state there) and the last position at which the
        initial state was left;
     4. STITCH: chunk i is confirmed when its start state equals the end state
        of chunk i-1 (chunk 0 starts in the initial state by definition).  The
        unconfirmed chunks get the correct start state and are walked again -
        all of them in parallel - until a round changes nothing.  A chunk
        whose guess was wrong usually still ENDS in the right state (the DFA
        converges inside the chunk), so a wrong guess does not cascade and the
        loop ends after one or two rounds; in the worst case (a DFA that never
        forgets, e.g. a parity counter) it degrades to one chunk per round and
        stays exact;
     5. reduce the per-chunk records in text order, and finish the styles
        that stop early with one short careful walk from the right chunk.

   Every answer is exact: nothing is derived from an unconfirmed chunk.
 */

#include "stream.h"
#include "walk.cuh"

#include <algorithm>
#include <type_traits>

namespace red_b200 {

namespace {

using namespace dev;

constexpr uint32_t kNoPos = 0xffffffffu;
constexpr uint32_t kLookback = 256;

struct ChunkRec {
  uint32_t startV;        // state the walk started from
  uint32_t endV;          // state after the chunk's last byte
  uint32_t firstAccOff;   // offset in chunk of the first accepting position, or kNoPos
  uint32_t firstAccV;
  uint32_t lastAccOff;    // ... of the last accepting position
  uint32_t lastAccV;
  uint32_t lastLeaveOff;  // last offset at which the initial state was left
  uint32_t dirty;         // needs (re)walking from startV
};

struct StreamSummary {
  uint64_t firstAccChunk, lastAccChunk, lastLeaveChunk;  // ~0 = none
  uint32_t firstAccOff, firstAccV, lastAccOff, lastAccV, lastLeaveOff;
  uint32_t finalV;
};

/* one chunk, from state v0; plain per-byte walk with event recording.
   Stops early once the state can never change again. */
template <class Tab, class ClsT>
__device__ __forceinline__ void walk_chunk(const ProgramDev &pd, const Tab &tab,
                                           const ClsT *cls, const uint8_t *p,
                                           uint32_t n, uint32_t v0, bool record,
                                           ChunkRec &rec) {
  uint32_t v = v0;
  uint32_t firstOff = kNoPos, firstV = 0, lastOff = kNoPos, lastV = 0, leaveOff = kNoPos;
  // p is 16-byte aligned (chunk sizes are multiples of 16; the text base is
  // required to be aligned); n may be short for the last chunk
  const uint4 *blk = reinterpret_cast<const uint4 *>(p);
  const uint32_t nblk = (n + 15) >> 4;
  for (uint32_t k = 0; k < nblk; ++k) {
    const uint4 q = __ldg(blk + k);
    const uint32_t hi = min(16u, n - k * 16);
    // fast pass
    uint32_t e = v, m = 0;
    bool left = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < hi) {
        const uint32_t b = byte_of(word_of(q, j >> 2), j & 3);
        const uint32_t ne = tab.next(e, cls[b]);
        left |= (e == pd.initV) && (ne != pd.initV);
        e = ne;
        m = max(m, ne);
      }
    }
    if (record && ((m >= pd.acceptV) || left)) {
      // careful pass over this block: something to record happened in it
      e = v;
#pragma unroll 1
      for (uint32_t j = 0; j < hi; ++j) {
        const uint32_t w = word_of(q, j >> 2);
        const uint32_t b = (w >> (8 * (j & 3))) & 0xffu;
        const uint32_t ne = tab.next(e, cls[b]);
        const uint32_t off = k * 16 + j;
        if ((e == pd.initV) && (ne != pd.initV))
          leaveOff = off;
        e = ne;
        if ((e >= pd.acceptV) && (e < pd.pureV)) {
          if (firstOff == kNoPos) {
            firstOff = off;
            firstV = e;
          }
          lastOff = off;
          lastV = e;
        }
      }
    }
    v = e;
    if (v >= pd.absV) {
      // absorbing accept (every later position accepts, same state) or pure
      // dead end (nothing can happen any more)
      if (record && (v < pd.pureV) && (n > 0)) {
        lastOff = n - 1;
        lastV = v;
      }
      break;
    }
  }
  rec.endV = v;
  if (record) {
    rec.firstAccOff = firstOff;
    rec.firstAccV = firstV;
    rec.lastAccOff = lastOff;
    rec.lastAccV = lastV;
    rec.lastLeaveOff = leaveOff;
  }
}

template <bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(kMaxThreads, 1)
stream_walk_kernel(const ProgramDev pd, const uint8_t *text, uint64_t len,
                   uint32_t chunkBytes, uint64_t nChunks, ChunkRec *recs,
                   int firstRound) {
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
  for (uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       c < nChunks; c += stride) {
    ChunkRec rec = recs[c];
    if (!firstRound && !rec.dirty)
      continue;
    const uint64_t begin = c * chunkBytes;
    const uint32_t n = static_cast<uint32_t>(min(static_cast<uint64_t>(chunkBytes), len - begin));
    if (firstRound) {
      // speculate: the state after the lookback window, entered from the
      // initial state (chunk 0 really starts in the initial state)
      uint32_t v = pd.initV;
      if (c > 0) {
        ChunkRec tmp;
        const uint32_t lb = static_cast<uint32_t>(min(static_cast<uint64_t>(kLookback), begin));
        walk_chunk(pd, tab, cls, text + begin - lb, lb, v, false, tmp);
        v = tmp.endV;
      }
      rec.startV = v;
    }
    walk_chunk(pd, tab, cls, text + begin, n, rec.startV, true, rec);
    rec.dirty = 0;
    recs[c] = rec;
  }
}

/* confirm chunk i against the end state of chunk i-1; hand the right start
   state to the ones that fail and count them */
__global__ void stream_verify_kernel(ChunkRec *recs, uint64_t nChunks,
                                     uint32_t initV,
                                     unsigned long long *mismatches) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  unsigned long long bad = 0;
  for (uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       c < nChunks; c += stride) {
    const uint32_t want = (c == 0) ? initV : recs[c - 1].endV;
    if (recs[c].startV != want) {
      recs[c].startV = want;
      recs[c].dirty = 1;
      ++bad;
    }
  }
  bad = __reduce_add_sync(0xffffffffu, static_cast<unsigned>(bad));
  if (((threadIdx.x & 31) == 0) && bad)
    atomicAdd(mismatches, bad);
}

/* in text order: first/last chunk with an accepting position, last chunk in
   which the initial state was left, final state.  One block; every thread
   takes a contiguous run of chunks, thread 0 combines. */
__global__ void __launch_bounds__(1024)
stream_reduce_kernel(const ChunkRec *recs, uint64_t nChunks, StreamSummary *out) {
  __shared__ unsigned long long sFirst[1024], sLast[1024], sLeave[1024];
  const uint64_t per = (nChunks + blockDim.x - 1) / blockDim.x;
  const uint64_t lo = min(nChunks, per * threadIdx.x);
  const uint64_t hi = min(nChunks, lo + per);
  unsigned long long f = ~0ull, l = ~0ull, lv = ~0ull;
  for (uint64_t c = lo; c < hi; ++c) {
    if (recs[c].firstAccOff != kNoPos) {
      if (f == ~0ull)
        f = c;
      l = c;
    }
    if (recs[c].lastLeaveOff != kNoPos)
      lv = c;
  }
  sFirst[threadIdx.x] = f;
  sLast[threadIdx.x] = l;
  sLeave[threadIdx.x] = lv;
  __syncthreads();
  if (threadIdx.x == 0) {
    f = l = lv = ~0ull;
    for (unsigned t = 0; t < blockDim.x; ++t) {
      if ((f == ~0ull) && (sFirst[t] != ~0ull))
        f = sFirst[t];
      if (sLast[t] != ~0ull)
        l = sLast[t];
      if (sLeave[t] != ~0ull)
        lv = sLeave[t];
    }
    StreamSummary s;
    s.firstAccChunk = f;
    s.lastAccChunk = l;
    s.lastLeaveChunk = lv;
    s.firstAccOff = (f != ~0ull) ? recs[f].firstAccOff : kNoPos;
    s.firstAccV = (f != ~0ull) ? recs[f].firstAccV : 0;
    s.lastAccOff = (l != ~0ull) ? recs[l].lastAccOff : kNoPos;
    s.lastAccV = (l != ~0ull) ? recs[l].lastAccV : 0;
    s.lastLeaveOff = (lv != ~0ull) ? recs[lv].lastLeaveOff : kNoPos;
    s.finalV = nChunks ? recs[nChunks - 1].endV : 0;
    *out = s;
  }
}

struct TailOut {
  int32_t result;
  uint32_t accepted;     // an accepting position was found
  uint64_t end;          // one past the last accepting position of the run
  uint64_t lastLeave;    // last position where the initial state was left, or ~0
};

/* First / Tangent (and Instant): the reference keeps walking after the first
   accepting position while the result stays the same / while it accepts
   (Matcher.h:455-476).  One thread, from the start of the chunk that holds the
   first accepting position, with the per-byte decisions of walk.cuh. */
template <bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(kMaxThreads, 1)
stream_tail_kernel(const ProgramDev pd, const uint8_t *text, uint64_t len,
                   uint64_t from, uint32_t v0, int style, TailOut *out) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  if (threadIdx.x != 0)
    return;
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes
                     : static_cast<const unsigned char *>(pd.table)};
  uint32_t v = v0;
  uint32_t accV = kNone;
  uint64_t end = 0, leave = ~0ull;
  bool absorbed = false;
  for (uint64_t pos = from; pos < len; ++pos) {
    const uint32_t nv = tab.next(v, cls[__ldg(text + pos)]);
    if ((v == pd.initV) && (nv != pd.initV))
      leave = pos;
    v = nv;
    if (v >= pd.pureV)
      break;
    if (v >= pd.acceptV) {
      if ((style == RED_B200_STY_FIRST) && (accV != kNone) && (v != accV) &&
          (pd.results[state_index(pd, v)] != pd.results[state_index(pd, accV)]))
        break;
      accV = v;
      end = pos + 1;
      if (style == RED_B200_STY_INSTANT)
        break;
      if (v >= pd.absV) {
        absorbed = true;
        break;
      }
    }
    else if (accV != kNone)
      break;  // First/Tangent: the accepting run is over
  }
  TailOut o;
  o.accepted = (accV != kNone);
  o.result = o.accepted ? pd.results[state_index(pd, accV)] : 0;
  o.end = absorbed ? len : end;
  o.lastLeave = leave;
  *out = o;
}

template <bool SMEM, int K, class ClsT>
int run_stream(const ProgramDev &pd, const LaunchGeometry &geo,
               const uint8_t *dText, uint64_t len, int style,
               red_b200_outcome *out, cudaStream_t stream, int *launches,
               std::string &why) {
  auto fail = [&](cudaError_t e, const char *what) {
    why = std::string(what) + ": " + cudaGetErrorString(e);
    return RED_B200_ECUDA;
  };
#define TRY(expr)                                \
  do {                                           \
    cudaError_t e_ = (expr);                     \
    if (e_ != cudaSuccess)                       \
      return fail(e_, #expr);                    \
  } while (0)

  const size_t smem = kClsBytes + (SMEM ? pd.tableBytes : 0);
  auto walk = stream_walk_kernel<SMEM, K, ClsT>;
  auto tail = stream_tail_kernel<SMEM, K, ClsT>;
  TRY(cudaFuncSetAttribute(walk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(tail, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));

  // chunk size: ~4 chunks per lane of the whole chip, 1 KB .. 16 KB, 16-byte multiple
  const uint64_t lanes = static_cast<uint64_t>(geo.smCount) * kMaxThreads;
  uint64_t chunk = (len + lanes * 4 - 1) / (lanes * 4);
  chunk = std::max<uint64_t>(1024, std::min<uint64_t>(16384, chunk));
  chunk = (chunk + 15) & ~uint64_t(15);
  const uint64_t nChunks = (len + chunk - 1) / chunk;

  ChunkRec *recs = nullptr;
  unsigned long long *counter = nullptr;
  StreamSummary *dSum = nullptr;
  TailOut *dTail = nullptr;
  TRY(cudaMallocAsync(&recs, nChunks * sizeof(ChunkRec), stream));
  TRY(cudaMallocAsync(&counter, sizeof(unsigned long long), stream));
  TRY(cudaMallocAsync(&dSum, sizeof(StreamSummary), stream));
  TRY(cudaMallocAsync(&dTail, sizeof(TailOut), stream));
  TRY(cudaMemsetAsync(recs, 0, nChunks * sizeof(ChunkRec), stream));

  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount, (nChunks + kMaxThreads - 1) / kMaxThreads));
  const unsigned vgrid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount * 4, (nChunks + 255) / 256));
  walk<<<grid, kMaxThreads, smem, stream>>>(pd, dText, len, static_cast<uint32_t>(chunk), nChunks, recs, 1);
  ++*launches;
  for (uint64_t round = 0;; ++round) {
    unsigned long long bad = 0;
    TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    stream_verify_kernel<<<vgrid, 256, 0, stream>>>(recs, nChunks, pd.initV, counter);
    ++*launches;
    TRY(cudaMemcpyAsync(&bad, counter, sizeof(bad), cudaMemcpyDeviceToHost, stream));
    TRY(cudaStreamSynchronize(stream));
    if (bad == 0)
      break;
    if (round > nChunks) {
      why = "stream stitching did not converge";
      return RED_B200_EINTERNAL;
    }
    walk<<<grid, kMaxThreads, smem, stream>>>(pd, dText, len, static_cast<uint32_t>(chunk), nChunks, recs, 0);
    ++*launches;
  }
  stream_reduce_kernel<<<1, 1024, 0, stream>>>(recs, nChunks, dSum);
  ++*launches;
  StreamSummary sum;
  TRY(cudaMemcpyAsync(&sum, dSum, sizeof(sum), cudaMemcpyDeviceToHost, stream));
  TRY(cudaStreamSynchronize(stream));

  auto resultOf = [&](uint32_t v, int32_t &r) -> int {
    uint32_t idx = v / pd.vPerState;
    TRY(cudaMemcpyAsync(&r, pd.results + idx, 4, cudaMemcpyDeviceToHost, stream));
    TRY(cudaStreamSynchronize(stream));
    return RED_B200_OK;
  };
  const uint64_t none = ~0ull;
  const uint64_t lastLeave = (sum.lastLeaveChunk != none)
      ? sum.lastLeaveChunk * chunk + sum.lastLeaveOff : 0;
  out->result = 0;
  out->start = 0;
  out->end = 0;
  int rc = RED_B200_OK;

  if (style == RED_B200_STY_FULL) {
    // the final state decides (a pure dead end carries result 0)
    const bool acc = (sum.finalV >= pd.acceptV) && (sum.finalV < pd.pureV);
    if (acc) {
      int32_t r = 0;
      if ((rc = resultOf(sum.finalV, r)) != RED_B200_OK)
        return rc;
      out->result = r;
      out->start = lastLeave;
      out->end = len;
    }
  }
  else if (style == RED_B200_STY_LAST) {
    if (sum.lastAccChunk != none) {
      int32_t r = 0;
      if ((rc = resultOf(sum.lastAccV, r)) != RED_B200_OK)
        return rc;
      out->result = r;
      out->start = lastLeave;
      out->end = sum.lastAccChunk * chunk + sum.lastAccOff + 1;
    }
  }
  else if (sum.firstAccChunk != none) {
    // Instant / First / Tangent: careful walk from the chunk with the first
    // accepting position
    const uint64_t c = sum.firstAccChunk;
    ChunkRec rec;
    TRY(cudaMemcpyAsync(&rec, recs + c, sizeof(rec), cudaMemcpyDeviceToHost, stream));
    TRY(cudaStreamSynchronize(stream));
    tail<<<1, kMaxThreads, smem, stream>>>(pd, dText, len, c * chunk, rec.startV, style, dTail);
    ++*launches;
    TailOut t;
    TRY(cudaMemcpyAsync(&t, dTail, sizeof(t), cudaMemcpyDeviceToHost, stream));
    TRY(cudaStreamSynchronize(stream));
    if (t.accepted) {
      out->result = t.result;
      out->end = t.end;
      // the initial state was last left either inside the tail walk or in an
      // earlier chunk
      uint64_t leave = t.lastLeave;
      if ((leave == none) && (c > 0)) {
        // the latest chunk in front of c in which the initial state was left
        stream_reduce_kernel<<<1, 1024, 0, stream>>>(recs, c, dSum);
        ++*launches;
        StreamSummary before;
        TRY(cudaMemcpyAsync(&before, dSum, sizeof(before), cudaMemcpyDeviceToHost, stream));
        TRY(cudaStreamSynchronize(stream));
        if (before.lastLeaveChunk != none)
          leave = before.lastLeaveChunk * chunk + before.lastLeaveOff;
      }
      if (leave == none)
        leave = 0;
      out->start = leave;
    }
  }
  // an empty text: the initial state's result (Matcher.h:429, loop not entered)
  if ((len == 0) && (pd.initResult > 0)) {
    out->result = pd.initResult;
    out->start = 0;
    out->end = 0;
  }
  TRY(cudaFreeAsync(recs, stream));
  TRY(cudaFreeAsync(counter, stream));
  TRY(cudaFreeAsync(dSum, stream));
  TRY(cudaFreeAsync(dTail, stream));
#undef TRY
  return RED_B200_OK;
}

} // anonymous

int match_stream(const ProgramDev &pd, bool smemTable, const DeviceImage &,
                 const LaunchGeometry &geo, const uint8_t *dText, uint64_t len,
                 int style, red_b200_outcome *out, cudaStream_t stream,
                 int *launches, std::string &why) {
  if (len == 0) {
    out->result = pd.initResult > 0 ? pd.initResult : 0;
    out->start = 0;
    out->end = 0;
    return RED_B200_OK;
  }
  if (reinterpret_cast<uintptr_t>(dText) & 15) {
    why = "stream text must be 16-byte aligned";
    return RED_B200_EAPI;
  }
  const bool cls8 = pd.clsMaxOffset < 256;
  if (smemTable) {
    if (pd.unitShift == 0)
      return cls8 ? run_stream<true, 1, uint8_t>(pd, geo, dText, len, style, out, stream, launches, why)
                  : run_stream<true, 1, uint16_t>(pd, geo, dText, len, style, out, stream, launches, why);
    if (pd.unitShift == 1)
      return cls8 ? run_stream<true, 2, uint8_t>(pd, geo, dText, len, style, out, stream, launches, why)
                  : run_stream<true, 2, uint16_t>(pd, geo, dText, len, style, out, stream, launches, why);
    if (pd.unitShift == 8)
      return run_stream<true, 9, uint16_t>(pd, geo, dText, len, style, out, stream, launches, why);
    why = "unsupported table layout";
    return RED_B200_EINTERNAL;
  }
  return cls8 ? run_stream<false, 0, uint8_t>(pd, geo, dText, len, style, out, stream, launches, why)
              : run_stream<false, 0, uint16_t>(pd, geo, dText, len, style, out, stream, launches, why);
}

} // namespace red_b200
