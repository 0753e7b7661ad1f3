/* stream_match.cu - match()/check() of ONE long byte stream
   (reference: matchCore over a single text, include/Matcher.h:413-495).

   A DFA walk is a chain of dependent lookups, so one stream has no
   parallelism unless the start state of a piece is known before the pieces
   in front of it are done.  What is done here:

     1. the text is taken in WINDOWS of growing size (32 MB, then x8): the
        state at the start of a window is known exactly from the window in
        front of it, and a walk that is decided early (Instant/First/Tangent
        on the first match, any style in a state that can never be left) ends
        after the window it was decided in, as the reference's does;
     2. a window is cut into chunks, one lane per chunk (1-16 KB each);
     3. every chunk is walked from a small SET OF CANDIDATE start states that
        is sure to hold the true one:
          - a DFA of at most 8 states: all of them (the state map of the
            chunk);
          - otherwise: the lane looks back from its chunk for the last byte b
            behind which the DFA can only be in a handful of states whatever
            came before (the image of ALL states under b, precomputed at
            upload: in a ".*pattern" automaton any byte that occurs in no
            pattern leads everything to the initial state), takes those, and
            walks them up to the chunk's start;
        with the batch kernels' chunk engine (chunk_engine.cuh) in style Last:
        the record of a (chunk, candidate) is its end state, its last accepting
        position with the state there and, when asked for, the last position
        at which the initial state was left;
     4. STITCH on the device: a chunk's candidates map to the next chunk's by
        their end states; these maps (eight 4-bit entries) are composed within
        a warp's 32 consecutive chunks, the warps' maps by ONE thread block in
        a three-level scan, and the true candidate of every chunk comes back
        down the same way.  No host round trip, three small launches.
     5. a chunk without such a byte within its lookback (a large DFA that
        forgets nothing) has no candidates; what lies behind the first of
        those in a window is resolved by one warp, in order, walking where it
        must - exact for every DFA, a bounded number of launches, slow only
        where the DFA leaves no other way;
     6. the styles that stop at the first match get one careful walk from the
        start of the chunk that holds it; a run of accepting positions that
        outlasts that chunk (Tangent, First) is ended in parallel: every chunk
        behind it reports the first position at which the run would break.

   Every answer is exact.
 */

#include "stream.h"
#include "chunk_engine.cuh"
#include "hot_walk.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace red_b200 {

namespace {

using namespace dev;

constexpr uint32_t kNoPos = 0xffffffffu;
constexpr uint32_t kLookback = 256;
constexpr uint64_t kFirstWindow = 32ull << 20;
constexpr uint64_t kWindowGrowth = 8;

constexpr uint32_t kCand = DeviceImage::kResetMax;  // candidate start states per chunk
constexpr uint32_t kNoCand = 0xfu;

// one walk of a chunk from one candidate start state
struct CandRec {
  uint32_t startV, endV, accV, accEnd, lastLeaveOff;
};

// a chunk's second, third ... candidate, walked by stream_extra_kernel
struct ExtraWalk {
  uint64_t chunk;
  uint32_t k, startV;
};

struct ChunkRec {
  uint32_t startV;        // state the walk started from
  uint32_t endV;          // state after the chunk's last byte (or the state
                          // that can never be left, if the walk stopped there)
  uint32_t accV;          // last accepting state in the chunk, or kNone
  uint32_t accEnd;        // offset one past that position
  uint32_t lastLeaveOff;  // last offset at which the initial state was left, or kNoPos
  uint32_t dirty;         // needs (re)walking from startV
};

struct StreamSummary {
  unsigned long long bad;        // unconfirmed chunks (written by the verify kernel)
  unsigned long long firstBad;   // first of them, or ~0
  // reduce kernel, by atomics: smallest chunk index with an accept (~0 none),
  // largest index + 1 with an accept / with a departure from the initial state
  unsigned long long firstAccChunk, lastAccP1, lastLeaveP1;
  unsigned int done;             // blocks of the reduce kernel that have finished
  uint32_t firstAccStartV;
  uint32_t lastAccEnd, lastAccV, lastLeaveOff;
  uint32_t finalV;
  uint64_t reduced;              // chunks the summary covers (a confirmed prefix)
};

/* n bytes at p (16-byte aligned) from state wk.v, 16 bytes per step through
   the chunk engine */
template <int STYLE, bool WANT_START, class Eng, class ClsT>
__device__ __forceinline__ void walk_bytes(const ProgramDev &pd, const typename Eng::Fast &fast,
                                           const typename Eng::Tab &tab, const ClsT *cls,
                                           const uint8_t *p, uint32_t n,
                                           Walk<STYLE, WANT_START> &wk) {
  const uint4 *blk = reinterpret_cast<const uint4 *>(p);
  const uint32_t nblk = (n + 15) >> 4;
  bool active = nblk > 0;
  uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur;
  if (active)
    cur = __ldg(blk);
  uint32_t k = 0;
  // every block but a lane's last is whole: a loop without masks or votes up
  // to the shortest lane's last block, then the general one
  const uint32_t kMin = __reduce_min_sync(0xffffffffu, active ? nblk : 0xffffffffu);
  if (kMin != 0xffffffffu) {
    for (; k + 1 < kMin; ++k) {
      if (active)
        nxt = __ldg(blk + k + 1);
      const bool stopped = process_chunk<STYLE, WANT_START>(pd, fast, tab, cls, cur,
                                                            active ? 0xffffu : 0u, k * 16, wk,
                                                            active, true);
      if (stopped)
        active = false;
      cur = nxt;
    }
  }
  while (__any_sync(0xffffffffu, active)) {
    uint32_t hi = 0;
    bool last = true;
    if (active) {
      const uint32_t rem = n - k * 16;
      hi = min(16u, rem);
      last = rem <= 16;
      if (!last)
        nxt = __ldg(blk + k + 1);
    }
    const uint32_t mask = 0xffffu >> (16 - hi);
    const bool stopped = process_chunk<STYLE, WANT_START>(pd, fast, tab, cls, cur, mask,
                                                          k * 16, wk, active);
    if (stopped || last)
      active = false;
    ++k;
    cur = nxt;
  }
}

/* text/len: the window; v0: the (exact) state at its start.  Chunk c is walked
   from each of its candidate start states: cands[c * kCand + k], nCand[c] of
   them (0: no candidates, see the file comment). */
template <bool SMEM, int K, class ClsT, bool TRACK>
__global__ void __launch_bounds__(kMaxThreads, 1)
stream_cand_kernel(const ProgramDev pd, const uint8_t *text, uint64_t len,
                   uint32_t chunkBytes, uint64_t nChunks, CandRec *cands, uint8_t *nCand,
                   uint32_t v0, uint32_t lookback, ExtraWalk *extras, unsigned long long *nExtras,
                   const bool walkFirst) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  using Eng = Engine<SMEM, K, ClsT>;
  unsigned char *smem = Eng::stage(pd, smemRaw);
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const typename Eng::Tab tab = Eng::table(pd, smem);
  const typename Eng::Fast fast = Eng::fast(pd, smem);

  // whole warps stay in the loop together (walk_bytes votes)
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t first = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t rounds = (nChunks + stride - 1) / stride;
  const bool allStates = pd.stateCount <= kCand;
  for (uint64_t r = 0; r < rounds; ++r) {
    const uint64_t c = first + r * stride;
    const bool mine = c < nChunks;
    const uint64_t begin = mine ? c * chunkBytes : 0;
    const uint32_t n = mine ? static_cast<uint32_t>(min(static_cast<uint64_t>(chunkBytes), len - begin)) : 0;
    uint32_t startV[kCand];
    uint32_t cnt = 0;
    if (mine) {
      if (c == 0) {
        startV[0] = v0;
        cnt = 1;
      }
      else if (allStates) {
        for (uint32_t k = 0; k < pd.stateCount; ++k)
          startV[k] = k * pd.vPerState;
        cnt = pd.stateCount;
      }
      else {
        // a byte in the lookback behind which only a few states are possible;
        // from each of them up to the chunk's start.  Not the nearest one:
        // with some text between it and the chunk the candidates have run
        // together (states that differ only in what lies far behind them
        // agree again after a byte or two), and every candidate that is left
        // is a walk of the whole chunk.  (allStates: the loop is not entered.)
        const uint32_t preferred = min(lookback, 48u);
        for (uint32_t step = 0; (step < 2 * lookback) && (cnt == 0); ++step) {
          // preferred, preferred + 1, ... lookback, then preferred - 1 ... 1
          const uint32_t back = (step <= lookback - preferred) ? preferred + step
                                                               : preferred - (step - (lookback - preferred));
          if ((back == 0) || (back > lookback))
            continue;
          const uint32_t byte = __ldg(text + begin - back);
          const uint32_t have = __ldg(pd.resetCount + byte);
          if (have == 0)
            continue;
          // all of them in step (eight independent lookups per byte; a slot
          // beyond `have` repeats the last state), duplicates dropped at the end
          uint32_t v[kCand];
#pragma unroll
          for (uint32_t k = 0; k < kCand; ++k)
            v[k] = __ldg(pd.resetStates + byte * kCand + min(k, have - 1));
          for (uint32_t i = back - 1; i > 0; --i) {
            const uint32_t co = cls[__ldg(text + begin - i)];
#pragma unroll
            for (uint32_t k = 0; k < kCand; ++k)
              v[k] = tab.next(v[k], co);
          }
#pragma unroll
          for (uint32_t k = 0; k < kCand; ++k) {
            bool dup = false;
#pragma unroll
            for (uint32_t j = 0; j < k; ++j)
              dup |= v[j] == v[k];
            if (!dup)
              startV[cnt++] = v[k];
          }
        }
      }
    }
    // the first candidate is walked here; the others - few chunks have any, but
    // most warps have such a chunk - go on a list for a launch of their own,
    // so that no warp waits for one lane's second walk
    for (uint32_t k = 1; k < cnt; ++k) {
      ExtraWalk e;
      e.chunk = c;
      e.k = k;
      e.startV = startV[k];
      extras[atomicAdd(nExtras, 1ull)] = e;
    }
    if (!walkFirst) {
      // the first candidate's walk is stream_hot_kernel's: only its start state
      if (cnt > 0)
        cands[c * kCand].startV = startV[0];
    }
    else {
      const uint32_t k = 0;
      const bool go = k < cnt;
      // TRACK: also the last exit from the initial state (three more
      // instructions per step: only the lazy start search asks for it)
      Walk<STY_LAST, TRACK> wk;
      wk.begin(pd);
      wk.v = go ? startV[k] : pd.initV;
      wk.mstart = kNoPos;
      walk_bytes<STY_LAST, TRACK, Eng>(pd, fast, tab, cls, text + begin, go ? n : 0, wk);
      if (go) {
        CandRec rec;
        rec.startV = startV[k];
        rec.endV = wk.v;
        rec.accV = wk.accV;
        rec.accEnd = wk.absorbed ? n : wk.mend;
        rec.lastLeaveOff = TRACK ? wk.mstart : kNoPos;
        cands[c * kCand + k] = rec;
      }
    }
    if (mine)
      nCand[c] = static_cast<uint8_t>(cnt);
  }
}

/* The first candidate's walk of every chunk on the HOT IMAGE (program.h), the
   text staged through shared memory.

   What bounded the walk above (ncu: the LSU data pipe, 91 %): two lookups per
   byte, and a 16-byte text load per lane and step that is a wavefront of its
   own - the lanes of a warp are a whole chunk apart.  Here a step is one
   lookup in the hot image (kernels_hot.cu's, from the same profile), and the
   text comes in stages of 64 bytes per lane: the warp copies the 32 rows of a
   stage together (cp.async, four lanes per row, so a request covers whole
   64-byte pieces) into rows of 80 bytes - an odd number of 16-byte units, so
   that the LDS.128 of eight consecutive lanes hit eight bank groups - two
   stages per warp, the next one in flight while the current one is walked.
   Lanes that leave the hot set go on on the full table, byte by byte, until
   they are back (hot_careful).  Style Last without start tracking, as the
   kernel above: end state, last accepting position and state. */
constexpr uint32_t kStageBytes = 64;
constexpr uint32_t kStagePitch = 80;
constexpr uint32_t kStageBuf = 32 * kStagePitch;
constexpr uint32_t kHotStreamMaxWarps = 28;

// shared-memory layout of stream_hot_kernel (byte offsets behind the hot table)
struct HotStreamParams {
  uint32_t oC2v, oCls, oCodeOf, oStage, oMirror;
  uint32_t codeOfBytes;  // 0: codeOf stays in global memory
  uint32_t mirrorBytes;  // leading bytes of the full table kept in shared memory
};

/* Where a lane that left the hot set walks: the full table, its leading rows
   (the states nearest the initial one) in shared memory, the rest in global
   memory; the class map and - if it fits - the state-to-code map in shared
   memory too.  Nearly every chunk has SOME lane of the warp out of the hot set
   for a byte or two (a 255-state hot set holds 99 % of the entries of a
   1191-state union, not all), and a byte on a table in L2 is two round trips
   with the warp waiting. */
// a step on the stream's image: entries of W bytes (hot_walk.cuh's hot_next is the W = 1 form)
template <int STRIDE, int W>
__device__ __forceinline__ uint32_t stream_next(const HotCtx &hx, uint32_t c, uint32_t word, int jj) {
  if (W == 1)
    return hot_next<STRIDE>(hx, c, word, jj);
  return lds16(c * STRIDE + hx.tab + 2 * byte_of(word, jj));
}
template <int STRIDE, int W>
__device__ __forceinline__ uint32_t stream_next_byte(const HotCtx &hx, uint32_t c, uint32_t byte) {
  if (W == 1)
    return hot_next_byte<STRIDE>(hx, c, byte);
  return lds16(c * STRIDE + hx.tab + 2 * byte);
}

struct StreamCold {
  const unsigned char *base;
  uint32_t shift, wide;
  uint32_t mirA, mirBytes;
  uint32_t clsA;
  uint32_t codeOfA;          // 0: not in shared memory
  const uint8_t *codeOfG;
  uint32_t codeWide;         // 1: two-byte codes
  __device__ __forceinline__ uint32_t next(uint32_t v, uint32_t co) const {
    const uint32_t off = (v << shift) + co;
    if (off < mirBytes)
      return wide ? lds32(mirA + off) : lds16(mirA + off);
    return wide ? __ldg(reinterpret_cast<const uint32_t *>(base + off))
                : static_cast<uint32_t>(__ldg(reinterpret_cast<const uint16_t *>(base + off)));
  }
  __device__ __forceinline__ uint32_t code_of(uint32_t idx) const {
    if (codeWide)
      return codeOfA ? lds16(codeOfA + 2 * idx)
                     : static_cast<uint32_t>(__ldg(reinterpret_cast<const uint16_t *>(codeOfG) + idx));
    return codeOfA ? lds8(codeOfA + idx) : static_cast<uint32_t>(__ldg(codeOfG + idx));
  }
};

/* Bytes [j, hi) of the chunk at chunkAddr one by one, style Last: on the hot
   table while the state is hot, on the full table while it is not.  c: the
   code in front of byte j in, behind the last byte taken out (hx.sink: the
   state is in wk.v).  True = the walk is over (a state that is never left). */
template <int STRIDE, int W>
__device__ __noinline__ bool stream_careful(const ProgramDev &pd, const HotCtx &hx, const StreamCold &sc,
                                               const uint32_t chunkAddr, uint32_t j, const uint32_t hi,
                                               const uint32_t pos0, uint32_t &c, Walk<STY_LAST, false> &wk) {
  constexpr bool kAsciiRows = STRIDE == 128 * W + 4;
#pragma unroll 1
  for (; j < hi; ++j) {
    const uint32_t byte = lds8(chunkAddr + j);
    bool stopped = false;
    if (c == hx.sink) {
      stopped = wk.step(pd, sc, lds16(sc.clsA + 2 * byte), pos0 + j);
      if (!stopped)
        c = sc.code_of(state_index(pd, wk.v));
    }
    else {
      const uint32_t nc = (kAsciiRows && (byte >= 128)) ? hx.sink : stream_next_byte<STRIDE, W>(hx, c, byte);
      if (nc >= hx.cLimit) {
        if (nc == hx.sink) {
          // leaves the hot set: this byte again, on the full table
          wk.v = lds32(hx.c2v + 4 * c);
          c = hx.sink;
          --j;
          continue;
        }
        if (nc >= hx.cPure)
          stopped = true;
        else {
          wk.accV = lds32(hx.c2v + 4 * nc);
          wk.mend = pos0 + j + 1;
          if (nc >= hx.cAbs) {
            wk.absorbed = true;
            stopped = true;
          }
        }
      }
      c = nc;
    }
    if (stopped)
      return true;
  }
  return false;
}

// (a copy goes through the call, so that wk and the code themselves can stay in registers)
template <int STRIDE, int W>
__device__ __forceinline__ bool stream_careful_call(const ProgramDev &pd, const HotCtx &hx, const StreamCold &sc,
                                                    const uint32_t chunkAddr, const uint32_t j, const uint32_t hi,
                                                    const uint32_t pos0, uint32_t &c, Walk<STY_LAST, false> &wk) {
  Walk<STY_LAST, false> tmp = wk;
  uint32_t cc = c;
  const bool stopped = stream_careful<STRIDE, W>(pd, hx, sc, chunkAddr, j, hi, pos0, cc, tmp);
  wk = tmp;
  c = cc;
  return stopped;
}

/* The chunk a lane noted last: its bytes again, from global memory, for the
   last accepting position in it (no state >= cAbs was entered there). */
template <int STRIDE, int W>
__device__ __forceinline__ void stream_resolve(const HotCtx &hx, const uint8_t *chunkText, const uint32_t hi,
                                               const uint32_t pos0, uint32_t c, Walk<STY_LAST, false> &wk) {
  const uint4 w = __ldg(reinterpret_cast<const uint4 *>(chunkText));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t word = (q == 0) ? w.x : (q == 1) ? w.y : (q == 2) ? w.z : w.w;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const uint32_t idx = static_cast<uint32_t>(4 * q + jj);
      if (idx < hi) {
        c = stream_next<STRIDE, W>(hx, c, word, jj);
        if (c >= hx.cLimit) {
          wk.accV = lds32(hx.c2v + 4 * c);
          wk.mend = pos0 + idx + 1;
        }
      }
    }
  }
}

/* One 16-byte chunk, style Last.  The fast pass keeps the largest code of the
   chunk.  An accepting state entered in it needs no attention now: of all the
   accepting positions of a lane's walk only the LAST one goes into the record,
   so the lane just remembers where the latest such chunk began and in which
   state (flagPos, flagCode) and looks into it once, when its walk is over
   (stream_resolve).  Only a state that is never left, the way out of the hot
   set, or a byte >= 128 on 128-column rows is dealt with on the spot. */
template <bool FULL16, int STRIDE, int W>
__device__ __forceinline__ bool stream_chunk_body(const ProgramDev &pd, const HotCtx &hx, const StreamCold &sc,
                                                  const uint4 &w, const uint32_t chunkAddr, const uint32_t hi,
                                                  const uint32_t pos0, uint32_t &code, Walk<STY_LAST, false> &wk,
                                                  uint32_t &flagPos, uint32_t &flagCode,
                                                  const uint8_t *laneText, const uint32_t laneBytes) {
  bool careful = code == hx.sink;
  uint32_t c = code, m = 0;
  if (!careful) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t word = (q == 0) ? w.x : (q == 1) ? w.y : (q == 2) ? w.z : w.w;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        if (FULL16 || (static_cast<uint32_t>(4 * q + jj) < hi))
          c = stream_next<STRIDE, W>(hx, c, word, jj);
        m = max(m, c);
      }
    }
    if ((STRIDE == 128 * W + 4) && (((w.x | w.y | w.z | w.w) & 0x80808080u) != 0))
      careful = true;
    else if (m < hx.cAbs) {
      if (m >= hx.cLimit) {
        flagPos = pos0;
        flagCode = code;
      }
      code = c;
      return false;
    }
    else
      careful = true;
  }
  // what was noted before is settled first: an accepting position recorded
  // from here on lies behind it
  if (flagPos != kNoPos) {
    stream_resolve<STRIDE, W>(hx, laneText + flagPos, min(16u, laneBytes - flagPos), flagPos, flagCode, wk);
    flagPos = kNoPos;
  }
  return stream_careful_call<STRIDE, W>(pd, hx, sc, chunkAddr, 0, hi, pos0, code, wk);
}

// (the kernel's loop holds ONE copy of the whole-chunk form; a chunk's ragged
// end - once per 1-16 KB - goes through a call)
template <int STRIDE, int W>
__device__ __noinline__ bool stream_chunk_tail(const ProgramDev &pd, const HotCtx &hx, const StreamCold &sc,
                                               const uint4 &w, const uint32_t chunkAddr, const uint32_t hi,
                                               const uint32_t pos0, uint32_t &code, Walk<STY_LAST, false> &wk,
                                               uint32_t &flagPos, uint32_t &flagCode,
                                               const uint8_t *laneText, const uint32_t laneBytes) {
  return stream_chunk_body<false, STRIDE, W>(pd, hx, sc, w, chunkAddr, hi, pos0, code, wk, flagPos, flagCode,
                                          laneText, laneBytes);
}
template <int STRIDE, int W>
__device__ __forceinline__ bool stream_chunk_tail_call(const ProgramDev &pd, const HotCtx &hx, const StreamCold &sc,
                                                       const uint4 &w, const uint32_t chunkAddr, const uint32_t hi,
                                                       const uint32_t pos0, uint32_t &code, Walk<STY_LAST, false> &wk,
                                                       uint32_t &flagPos, uint32_t &flagCode,
                                                       const uint8_t *laneText, const uint32_t laneBytes) {
  Walk<STY_LAST, false> tmp = wk;
  uint32_t cc = code, fp = flagPos, fc = flagCode;
  const uint4 ww = w;
  const bool stopped = stream_chunk_tail<STRIDE, W>(pd, hx, sc, ww, chunkAddr, hi, pos0, cc, tmp, fp, fc, laneText,
                                                 laneBytes);
  wk = tmp;
  code = cc;
  flagPos = fp;
  flagCode = fc;
  return stopped;
}

template <int STRIDE, int W>
__global__ void __launch_bounds__(kHotStreamMaxWarps * 32, 1)
stream_hot_kernel(const ProgramDev pd, const HotDev hd, const HotStreamParams sp, const uint8_t *text, uint64_t len,
                  uint32_t chunkBytes, uint64_t nChunks, CandRec *cands, const uint8_t *nCand) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  unsigned char *smem = smemRaw + ((256u - (smem_addr(smemRaw) & 255u)) & 255u);
  const uint32_t sBase = smem_addr(smem);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t nWarps = blockDim.x >> 5;
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(hd.table);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (uint32_t i = threadIdx.x; i < (hd.tableBytes >> 4); i += blockDim.x)
      dst[i] = __ldg(src + i);
    uint32_t *c2v = reinterpret_cast<uint32_t *>(smem + sp.oC2v);
    for (uint32_t i = threadIdx.x; i <= hd.nHot; i += blockDim.x)
      c2v[i] = __ldg(hd.codeToV + i);
    uint16_t *cls = reinterpret_cast<uint16_t *>(smem + sp.oCls);
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x)
      cls[i] = __ldg(pd.clsOffset + i);
    for (uint32_t i = threadIdx.x; i < sp.codeOfBytes; i += blockDim.x)
      smem[sp.oCodeOf + i] = (i < pd.stateCount * W) ? __ldg(hd.codeOf + i) : static_cast<uint8_t>(0);
    const uint4 *tsrc = reinterpret_cast<const uint4 *>(pd.table);
    uint4 *tdst = reinterpret_cast<uint4 *>(smem + sp.oMirror);
    for (uint32_t i = threadIdx.x; i < (sp.mirrorBytes >> 4); i += blockDim.x)
      tdst[i] = __ldg(tsrc + i);
  }
  __syncthreads();
  HotCtx hx;
  hx.tab = sBase;
  hx.c2v = sBase + sp.oC2v;
  hx.sink = hd.nHot;
  hx.cLimit = hd.cAccept;
  hx.cAbs = hd.cAbs;
  hx.cPure = hd.cPure;
  hx.wide = hd.wideTable;
  hx.codeOf = hd.codeOf;
  StreamCold sc;
  sc.base = static_cast<const unsigned char *>(pd.table);
  sc.wide = hd.wideTable;
  sc.shift = hd.wideTable ? 2 : pd.unitShift + 1;
  sc.mirA = sBase + sp.oMirror;
  sc.mirBytes = sp.mirrorBytes;
  sc.clsA = sBase + sp.oCls;
  sc.codeOfA = sp.codeOfBytes ? sBase + sp.oCodeOf : 0;
  sc.codeOfG = hd.codeOf;
  sc.codeWide = (W == 2) ? 1 : 0;

  const uint32_t bufA = sBase + sp.oStage + warp * 2 * kStageBuf;
  const uint32_t rowA = lane * kStagePitch;
  const uint32_t nStages = (chunkBytes + kStageBytes - 1) / kStageBytes;
  const uint64_t totalWarps = static_cast<uint64_t>(gridDim.x) * nWarps;
  for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * nWarps + warp; g * 32 < nChunks; g += totalWarps) {
    const uint64_t c0 = g * 32;
    const uint64_t c = c0 + lane;
    const bool mine = c < nChunks;
    const uint64_t begin = mine ? c * chunkBytes : 0;
    const uint32_t n = mine ? static_cast<uint32_t>(min(static_cast<uint64_t>(chunkBytes), len - begin)) : 0;
    const bool go = mine && (__ldg(nCand + c) != 0);
    const uint32_t startV = go ? cands[c * kCand].startV : pd.initV;
    Walk<STY_LAST, false> wk;
    wk.begin(pd);
    wk.v = startV;
    uint32_t code = sc.code_of(state_index(pd, startV));
    bool active = go && (n != 0);
    uint32_t flagPos = kNoPos, flagCode = 0;  // the last chunk in which an accepting state was entered

    // the rows this lane copies in every stage: row (lane % 8) + 8 i, piece lane / 8
    // (the eight lanes of a quarter warp write to eight rows: 80 bytes apart,
    // eight different bank groups; the four pieces of a row, one per quarter
    // warp, are one 64-byte request to L2)
    const uint32_t piece = (lane >> 3) * 16;
    auto issue = [&](uint32_t st, uint32_t buf) {
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t row = i * 8 + (lane & 7);
        const uint64_t rc = c0 + row;
        const uint64_t rBegin = rc * chunkBytes;
        const uint64_t rBytes = (rc < nChunks) ? min(static_cast<uint64_t>(chunkBytes), len - rBegin) : 0;
        const uint64_t off = static_cast<uint64_t>(st) * kStageBytes + piece;
        const bool valid = off < rBytes;  // (a piece that starts inside the text lies in its last 16-byte block at most)
        cp_async16(buf + row * kStagePitch + piece, text + (valid ? rBegin + off : 0), valid ? 16u : 0u);
      }
      cp_async_commit();
    };
    issue(0, bufA);
    for (uint32_t st = 0; st < nStages; ++st) {
      const uint32_t cur = bufA + (st & 1) * kStageBuf;
      const bool more = st + 1 < nStages;
      if (more)
        issue(st + 1, bufA + ((st + 1) & 1) * kStageBuf);
      // the text a few stages on into L2, so that the copies find it there
      if (active && ((st & 1) == 0)) {
        const uint64_t ahead = static_cast<uint64_t>(st + 6) * kStageBytes;
        if (ahead < n)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(text + begin + ahead));
      }
      if (more)
        cp_async_wait<1>();
      else
        cp_async_wait<0>();
      __syncwarp();
#pragma unroll 1
      for (uint32_t j = 0; j < kStageBytes / 16; ++j) {
        const uint32_t pos = st * kStageBytes + 16 * j;
        if (active && (pos < n)) {
          const uint32_t a = cur + rowA + 16 * j;
          const uint4 w = lds128(a);
          const uint32_t rem = n - pos;
          const bool stopped = (rem >= 16)
              ? stream_chunk_body<true, STRIDE, W>(pd, hx, sc, w, a, 16, pos, code, wk, flagPos, flagCode,
                                                   text + begin, n)
              : stream_chunk_tail_call<STRIDE, W>(pd, hx, sc, w, a, rem, pos, code, wk, flagPos, flagCode,
                                                  text + begin, n);
          if (stopped || (rem <= 16))
            active = false;
        }
      }
      __syncwarp();  // everybody is through with `cur` before the stage after next lands in it
      if (!__any_sync(kFullMask, active)) {
        cp_async_wait<0>();
        __syncwarp();
        break;
      }
    }
    if (go && (flagPos != kNoPos))
      stream_resolve<STRIDE, W>(hx, text + begin + flagPos, min(16u, n - flagPos), flagPos, flagCode, wk);
    if (go) {
      CandRec rec;
      rec.startV = startV;
      rec.endV = (code != hx.sink) ? lds32(hx.c2v + 4 * code) : wk.v;
      rec.accV = wk.accV;
      rec.accEnd = wk.absorbed ? n : wk.mend;
      rec.lastLeaveOff = kNoPos;
      cands[c * kCand] = rec;
    }
  }
}

/* the walks stream_cand_kernel put on the list, one lane each */
template <bool SMEM, int K, class ClsT, bool TRACK>
__global__ void __launch_bounds__(kMaxThreads, 1)
stream_extra_kernel(const ProgramDev pd, const uint8_t *text, uint64_t len, uint32_t chunkBytes,
                    CandRec *cands, const ExtraWalk *extras, const unsigned long long *nExtras) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const unsigned long long total = *nExtras;
  if (total == 0)
    return;
  using Eng = Engine<SMEM, K, ClsT>;
  unsigned char *smem = Eng::stage(pd, smemRaw);
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const typename Eng::Tab tab = Eng::table(pd, smem);
  const typename Eng::Fast fast = Eng::fast(pd, smem);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t first = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t rounds = (total + stride - 1) / stride;
  for (uint64_t r = 0; r < rounds; ++r) {
    const uint64_t item = first + r * stride;
    const bool mine = item < total;
    ExtraWalk e{};
    if (mine)
      e = extras[item];
    const uint64_t begin = e.chunk * chunkBytes;
    const uint32_t n = mine ? static_cast<uint32_t>(min(static_cast<uint64_t>(chunkBytes), len - begin)) : 0;
    Walk<STY_LAST, TRACK> wk;
    wk.begin(pd);
    wk.v = mine ? e.startV : pd.initV;
    wk.mstart = kNoPos;
    walk_bytes<STY_LAST, TRACK, Eng>(pd, fast, tab, cls, text + begin, n, wk);
    if (mine) {
      CandRec rec;
      rec.startV = e.startV;
      rec.endV = wk.v;
      rec.accV = wk.accV;
      rec.accEnd = wk.absorbed ? n : wk.mend;
      rec.lastLeaveOff = TRACK ? wk.mstart : kNoPos;
      cands[e.chunk * kCand + e.k] = rec;
    }
  }
}

// a map between the candidates of two chunks: eight 4-bit entries, kNoCand = no image
__device__ __forceinline__ uint32_t map_then(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll
  for (uint32_t k = 0; k < kCand; ++k) {
    const uint32_t mid = (a >> (4 * k)) & 0xf;
    const uint32_t to = (mid == kNoCand) ? kNoCand : ((b >> (4 * mid)) & 0xf);
    r |= to << (4 * k);
  }
  return r;
}
__device__ __forceinline__ uint32_t map_apply(uint32_t m, uint32_t idx) {
  return (idx == kNoCand) ? kNoCand : ((m >> (4 * idx)) & 0xf);
}
constexpr uint32_t kMapIdentity = 0x76543210u;

/* maps[c]: candidate k of chunk c -> the candidate of chunk c + 1 that starts in
   k's end state (kNoCand if none: chunk c + 1 has no candidates); segMaps[s]:
   the 32 maps of warp segment s composed.  One lane per chunk. */
__global__ void __launch_bounds__(256)
stream_maps_kernel(const CandRec *cands, const uint8_t *nCand, uint64_t nChunks,
                   uint32_t *maps, uint32_t *segMaps) {
  const uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t m = kMapIdentity;  // (past the last chunk: harmless)
  if (c < nChunks) {
    m = 0;
    const uint32_t mine = nCand[c];
    const uint32_t next = (c + 1 < nChunks) ? nCand[c + 1] : 0;
    for (uint32_t k = 0; k < kCand; ++k) {
      uint32_t to = kNoCand;
      if (k < mine) {
        const uint32_t endV = cands[c * kCand + k].endV;
        for (uint32_t j = 0; j < next; ++j)
          if (cands[(c + 1) * kCand + j].startV == endV)
            to = j;
        if (c + 1 >= nChunks)
          to = k;  // nothing behind the last chunk
      }
      m |= to << (4 * k);
    }
    maps[c] = m;
  }
  uint32_t seg = kMapIdentity;
  for (uint32_t i = 0; i < 32; ++i)
    seg = map_then(seg, __shfl_sync(0xffffffffu, m, i));
  if ((lane == 0) && (c < nChunks))
    segMaps[c >> 5] = seg;
}

/* ONE block: the candidate index in force at the start of every warp segment
   (segIn), from index 0 at the start of the window - thread maps, warp maps,
   a walk over the warp maps, and back down. */
__global__ void __launch_bounds__(1024)
stream_scan_kernel(const uint32_t *segMaps, uint64_t nSegs, uint8_t *segIn) {
  __shared__ uint32_t warpMap[32];
  __shared__ uint32_t warpIn[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t per = (nSegs + 1023) / 1024;
  const uint64_t lo = min(nSegs, threadIdx.x * per), hi = min(nSegs, lo + per);
  uint32_t mine = kMapIdentity;
  for (uint64_t sgm = lo; sgm < hi; ++sgm)
    mine = map_then(mine, segMaps[sgm]);
  uint32_t w = kMapIdentity;
  for (uint32_t i = 0; i < 32; ++i)
    w = map_then(w, __shfl_sync(0xffffffffu, mine, i));
  if (lane == 0)
    warpMap[warp] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t idx = 0;
    for (uint32_t i = 0; i < 32; ++i) {
      warpIn[i] = idx;
      idx = map_apply(warpMap[i], idx);
    }
  }
  __syncthreads();
  uint32_t idx = warpIn[warp];
  uint32_t myIn = idx;
  for (uint32_t i = 0; i < 32; ++i) {
    if (lane == i)
      myIn = idx;
    idx = map_apply(__shfl_sync(0xffffffffu, mine, i), idx);
  }
  idx = myIn;
  for (uint64_t sgm = lo; sgm < hi; ++sgm) {
    segIn[sgm] = static_cast<uint8_t>(idx);
    idx = map_apply(segMaps[sgm], idx);
  }
}

/* the true candidate of every chunk -> recs[c]; chunks behind one without
   candidates stay unresolved (dirty) and are counted in sum->bad */
__global__ void __launch_bounds__(256)
stream_apply_kernel(const CandRec *cands, const uint32_t *maps, const uint8_t *segIn,
                    uint64_t nChunks, ChunkRec *recs, StreamSummary *sum) {
  const uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t m = (c < nChunks) ? maps[c] : kMapIdentity;
  uint32_t idx = (c < nChunks) ? segIn[c >> 5] : 0;
  idx = __shfl_sync(0xffffffffu, idx, 0);
  uint32_t mineIdx = idx;
  for (uint32_t i = 0; i < 32; ++i) {
    if (lane == i)
      mineIdx = idx;
    idx = map_apply(__shfl_sync(0xffffffffu, m, i), idx);
  }
  bool bad = false;
  if (c < nChunks) {
    ChunkRec r{};
    if (mineIdx == kNoCand) {
      r.dirty = 1;
      bad = true;
    }
    else {
      const CandRec k = cands[c * kCand + mineIdx];
      r.startV = k.startV;
      r.endV = k.endV;
      r.accV = k.accV;
      r.accEnd = k.accEnd;
      r.lastLeaveOff = k.lastLeaveOff;
    }
    recs[c] = r;
  }
  const unsigned bads = __ballot_sync(0xffffffffu, bad);
  if (bads && (lane == 0)) {
    atomicAdd(&sum->bad, static_cast<unsigned long long>(__popc(bads)));
    atomicMin(&sum->firstBad, static_cast<unsigned long long>(c + __ffs(bads) - 1));
  }
}

/* The chunks from sum->firstBad on, in order, by one thread: the true start
   state is the end state of the chunk in front; a candidate that starts there
   is taken, else the chunk is walked here, byte by byte on the full table.
   Exact for every DFA; only a large automaton that forgets nothing (no byte
   narrows its state down) comes here for more than the odd chunk. */
template <bool TRACK>
__global__ void stream_repair_kernel(const ProgramDev pd, uint32_t wide, const uint8_t *text, uint64_t len,
                                     uint32_t chunkBytes, uint64_t nChunks, const CandRec *cands,
                                     const uint8_t *nCand, ChunkRec *recs, StreamSummary *sum) {
  if ((blockIdx.x != 0) || (threadIdx.x != 0))
    return;
  const AnyTable tab = any_table(pd, wide);
  for (uint64_t c = sum->firstBad; c < nChunks; ++c) {
    const uint32_t startV = recs[c - 1].endV;  // (chunk 0 always has its candidate)
    ChunkRec r{};
    r.startV = startV;
    bool found = false;
    for (uint32_t k = 0; (k < nCand[c]) && !found; ++k)
      if (cands[c * kCand + k].startV == startV) {
        const CandRec cr = cands[c * kCand + k];
        r.endV = cr.endV;
        r.accV = cr.accV;
        r.accEnd = cr.accEnd;
        r.lastLeaveOff = cr.lastLeaveOff;
        found = true;
      }
    if (!found) {
      const uint64_t begin = c * chunkBytes;
      const uint32_t n = static_cast<uint32_t>(min(static_cast<uint64_t>(chunkBytes), len - begin));
      Walk<STY_LAST, TRACK> wk;
      wk.begin(pd);
      wk.v = startV;
      wk.mstart = kNoPos;
      for (uint32_t i = 0; i < n; ++i)
        if (wk.step(pd, tab, __ldg(pd.clsOffset + __ldg(text + begin + i)), i))
          break;
      r.endV = wk.v;
      r.accV = wk.accV;
      r.accEnd = wk.absorbed ? n : wk.mend;
      r.lastLeaveOff = TRACK ? wk.mstart : kNoPos;
    }
    recs[c] = r;
  }
  sum->bad = 0;
  sum->firstBad = ~0ull;
}

/* over the confirmed prefix: first/last chunk with an accepting position, last
   chunk in which the initial state was left, final state.  Any grid; the
   block that finishes last fills in the per-chunk details.
   upTo = ~0: the prefix in front of sum->firstBad (after the verify kernel). */
__global__ void __launch_bounds__(256)
stream_reduce_kernel(const ChunkRec *recs, uint64_t nChunks, uint64_t upTo,
                     StreamSummary *sum) {
  const uint64_t n = (upTo != ~0ull) ? min(upTo, nChunks)
                                     : min(static_cast<uint64_t>(sum->firstBad), nChunks);
  unsigned long long f = ~0ull, l = 0, lv = 0;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < n; c += stride) {
    if (recs[c].accV != kNone) {
      f = min(f, static_cast<unsigned long long>(c));
      l = c + 1;
    }
    if (recs[c].lastLeaveOff != kNoPos)
      lv = c + 1;
  }
  __syncwarp();
  for (int d = 16; d > 0; d >>= 1) {
    f = min(f, __shfl_xor_sync(0xffffffffu, f, d));
    l = max(l, __shfl_xor_sync(0xffffffffu, l, d));
    lv = max(lv, __shfl_xor_sync(0xffffffffu, lv, d));
  }
  if ((threadIdx.x & 31) == 0) {
    if (f != ~0ull)
      atomicMin(&sum->firstAccChunk, f);
    if (l)
      atomicMax(&sum->lastAccP1, l);
    if (lv)
      atomicMax(&sum->lastLeaveP1, lv);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(&sum->done, 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      f = atomicMin(&sum->firstAccChunk, ~0ull);
      l = atomicMax(&sum->lastAccP1, 0ull);
      lv = atomicMax(&sum->lastLeaveP1, 0ull);
      sum->firstAccStartV = (f != ~0ull) ? recs[f].startV : 0;
      sum->lastAccEnd = l ? recs[l - 1].accEnd : 0;
      sum->lastAccV = l ? recs[l - 1].accV : 0;
      sum->lastLeaveOff = lv ? recs[lv - 1].lastLeaveOff : kNoPos;
      sum->finalV = n ? recs[n - 1].endV : 0;
      sum->reduced = n;
    }
  }
}

struct TailOut {
  int32_t result;
  uint32_t accepted;     // an accepting position was found
  uint64_t end;          // one past the last accepting position of the run
  uint64_t lastLeave;    // last position where the initial state was left, or ~0
  uint32_t open;         // the run of accepting positions was still going at `until`
  uint32_t accV;         // the last accepting state
};

/* First / Tangent (and Instant): the reference keeps walking after the first
   accepting position while the result stays the same / while it accepts
   (Matcher.h:455-476).  One thread, from the start of the chunk that holds the
   first accepting position, with the per-byte decisions of walk.cuh. */
template <bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(kMaxThreads, 1)
stream_tail_kernel(const ProgramDev pd, const uint8_t *text, uint64_t len,
                   uint64_t from, uint64_t until, uint32_t v0, int style, TailOut *out) {
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
  bool absorbed = false, broke = false;
  for (uint64_t pos = from; pos < until; ++pos) {
    const uint32_t nv = tab.next(v, cls[__ldg(text + pos)]);
    if ((v == pd.initV) && (nv != pd.initV))
      leave = pos;
    v = nv;
    if (v >= pd.pureV) {
      broke = true;
      break;
    }
    if (v >= pd.acceptV) {
      if ((style == RED_B200_STY_FIRST) && (accV != kNone) && (v != accV) &&
          (pd.results[state_index(pd, v)] != pd.results[state_index(pd, accV)])) {
        broke = true;
        break;
      }
      accV = v;
      end = pos + 1;
      if (style == RED_B200_STY_INSTANT) {
        broke = true;
        break;
      }
      if (v >= pd.absV) {
        absorbed = true;
        break;
      }
    }
    else if (accV != kNone)
    {
      broke = true;  // First/Tangent: the accepting run is over
      break;
    }
  }
  TailOut o;
  o.accepted = (accV != kNone);
  o.result = o.accepted ? pd.results[state_index(pd, accV)] : 0;
  o.end = absorbed ? len : end;
  o.lastLeave = leave;
  o.open = (o.accepted && !absorbed && !broke && (until < len)) ? 1 : 0;
  o.accV = accV;
  *out = o;
}

/* A run of accepting positions that outlasted the chunk of its first one
   (Tangent: while the state accepts; First: while it accepts with the same
   result, Matcher.h:455-476).  Every chunk of the window behind that chunk is
   walked from its true start state as if the run were still going, up to the
   first position that would end it. */
struct BreakRec {
  uint32_t kind;    // 0 the run goes through the chunk, 1 it breaks at off, 2 it is absorbed (runs to the end of the text)
  uint32_t off;
  uint32_t accV;    // the last accepting state in front of the break (or of the chunk's end)
  uint32_t leave;   // last offset (in front of the break) at which the initial state was left, or kNoPos
};

template <bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(kMaxThreads, 1)
stream_break_kernel(const ProgramDev pd, const uint8_t *text, uint64_t len, uint32_t chunkBytes,
                    uint64_t firstChunk, uint64_t nChunks, const ChunkRec *recs, int style,
                    int32_t runResult, BreakRec *out) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes : static_cast<const unsigned char *>(pd.table)};
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t c = firstChunk + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < nChunks;
       c += stride) {
    const uint64_t begin = c * chunkBytes;
    const uint32_t n = static_cast<uint32_t>(min(static_cast<uint64_t>(chunkBytes), len - begin));
    uint32_t v = recs[c].startV;
    BreakRec r;
    r.kind = 0;
    r.off = n;
    r.accV = v;
    r.leave = kNoPos;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t nv = tab.next(v, cls[__ldg(text + begin + i)]);
      const bool left = (v == pd.initV) && (nv != pd.initV);
      v = nv;
      bool ends = (v >= pd.pureV) || (v < pd.acceptV);
      if (!ends && (style == RED_B200_STY_FIRST) && (v != r.accV) &&
          (pd.results[state_index(pd, v)] != runResult))
        ends = true;
      if (left)
        r.leave = i;  // (the reference notes it before it looks at the new state)
      if (ends) {
        r.kind = 1;
        r.off = i;
        break;
      }
      r.accV = v;
      if (v >= pd.absV) {
        r.kind = 2;
        r.off = i;
        break;
      }
    }
    out[c] = r;
  }
}

/* ONE block: the first chunk in [firstChunk, nChunks) at which the run ends,
   and the last exit from the initial state in front of that point. */
struct BreakSummary {
  unsigned long long chunk;   // ~0: the run goes through the whole window
  uint32_t kind, off, accV;
  unsigned long long leaveChunk;  // chunk of the last exit from the initial state, ~0: none
  uint32_t leaveOff;
  uint32_t endV;              // accV at the window's end when the run goes on
};
__global__ void __launch_bounds__(1024)
stream_break_reduce_kernel(const BreakRec *recs, uint64_t firstChunk, uint64_t nChunks, BreakSummary *out) {
  __shared__ unsigned long long firstEnd, lastLeave;
  if (threadIdx.x == 0) {
    firstEnd = ~0ull;
    lastLeave = 0;
  }
  __syncthreads();
  unsigned long long mine = ~0ull;
  for (uint64_t c = firstChunk + threadIdx.x; c < nChunks; c += blockDim.x)
    if (recs[c].kind != 0) {
      mine = c;
      break;
    }
  if (mine != ~0ull)
    atomicMin(&firstEnd, mine);
  __syncthreads();
  const unsigned long long stop = firstEnd;
  const uint64_t upTo = (stop == ~0ull) ? nChunks : stop + 1;  // the breaking chunk's own exits count, too
  unsigned long long lv = 0;
  for (uint64_t c = firstChunk + threadIdx.x; c < upTo; c += blockDim.x)
    if (recs[c].leave != kNoPos)
      lv = c + 1;
  if (lv)
    atomicMax(&lastLeave, lv);
  __syncthreads();
  if (threadIdx.x == 0) {
    BreakSummary s;
    s.chunk = stop;
    s.kind = (stop != ~0ull) ? recs[stop].kind : 0;
    s.off = (stop != ~0ull) ? recs[stop].off : 0;
    s.accV = (stop != ~0ull) ? recs[stop].accV : 0;
    s.leaveChunk = lastLeave ? lastLeave - 1 : ~0ull;
    s.leaveOff = lastLeave ? recs[lastLeave - 1].leave : 0;
    s.endV = (nChunks > firstChunk) ? recs[nChunks - 1].accV : 0;
    *out = s;
  }
}

template <bool SMEM, int K, class ClsT>
int run_stream(const ProgramDev &pd, const HotDev *hot, const LaunchGeometry &geo,
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

  const size_t smem = kLanesSmemExtra + (SMEM ? pd.tableBytes : pd.hotBytes);
  auto walkPlain = stream_cand_kernel<SMEM, K, ClsT, false>;
  auto walkTrack = stream_cand_kernel<SMEM, K, ClsT, true>;
  auto extraPlain = stream_extra_kernel<SMEM, K, ClsT, false>;
  auto extraTrack = stream_extra_kernel<SMEM, K, ClsT, true>;
  auto tail = stream_tail_kernel<SMEM, K, ClsT>;
  auto brk = stream_break_kernel<SMEM, K, ClsT>;
  TRY(cudaFuncSetAttribute(walkPlain, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(walkTrack, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(tail, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(extraPlain, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(extraTrack, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(brk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));

  /* The main walk (style Last, no start tracking) on the hot image where the
     program has one that holds (nearly) every state the profiled text enters:
     stream_hot_kernel.  RED_B200_STREAM_HOT=0 keeps it out, =1 takes it
     whatever the profile's coverage. */
  bool useHot = (hot != nullptr) && (hot->table != nullptr) && (hot->coverage >= 990);
  if (const char *env = getenv("RED_B200_STREAM_HOT"))
    useHot = (hot != nullptr) && (hot->table != nullptr) && (atoi(env) != 0);
  HotStreamParams hsp{};
  uint32_t hotWarpsMax = 0, hotFixed = 0;
  if (useHot) {
    uint32_t at = (hot->tableBytes + 15u) & ~15u;
    hsp.oC2v = at;
    at += ((hot->nHot + 1) * 4 + 15u) & ~15u;
    hsp.oCls = at;
    at += 512;
    hsp.codeOfBytes = (pd.stateCount <= 8192) ? ((pd.stateCount * hot->entryBytes + 15u) & ~15u) : 0;
    hsp.oCodeOf = at;
    at += hsp.codeOfBytes;
    hsp.oStage = at;
    const uint32_t room = (static_cast<uint32_t>(geo.maxSmemOptin) > at + 256) ? geo.maxSmemOptin - at - 256 : 0;
    // staging for as many warps as keep the load/store pipe busy, the rest of
    // the shared memory for the leading rows of the full table
    hotWarpsMax = std::min<uint32_t>(kHotStreamMaxWarps, room / (2 * kStageBuf));
    const uint32_t wantWarps = 24;
    if ((hot->coverage < 1000) && (hotWarpsMax > wantWarps) && (room - wantWarps * 2 * kStageBuf < pd.tableBytes))
      hotWarpsMax = wantWarps;
    if (const char *env = getenv("RED_B200_STREAM_HOT_WARPS"))
      hotWarpsMax = std::max(1u, std::min<uint32_t>(hotWarpsMax, static_cast<uint32_t>(atoi(env))));
    hsp.oMirror = hsp.oStage + hotWarpsMax * 2 * kStageBuf;
    hsp.mirrorBytes = (hotWarpsMax >= 4)
        ? std::min<uint32_t>(pd.tableBytes & ~15u, (room - hotWarpsMax * 2 * kStageBuf) & ~15u) : 0;
    hotFixed = hsp.oMirror + hsp.mirrorBytes + 256;
  }
  if (hotWarpsMax < 4)
    useHot = false;
  auto hotWalk = stream_hot_kernel<132, 1>;
  if (useHot) {
    if (hot->entryBytes == 2)
      hotWalk = (hot->stride == 260) ? stream_hot_kernel<260, 2> : stream_hot_kernel<516, 2>;
    else
      hotWalk = (hot->stride == kHotStrideAscii) ? stream_hot_kernel<132, 1> : stream_hot_kernel<260, 1>;
  }
  if (useHot)
    TRY(cudaFuncSetAttribute(hotWalk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hotFixed)));

  // chunk size of a window: one chunk per lane of the whole chip if that
  // gives 1 KB .. 16 KB, a multiple of 16
  const uint64_t lanes = static_cast<uint64_t>(geo.smCount) * kMaxThreads;
  auto chunkFor = [&](uint64_t bytes) {
    uint64_t chunk = (bytes + lanes - 1) / lanes;
    chunk = std::max<uint64_t>(1024, std::min<uint64_t>(16384, chunk));
    return (chunk + 15) & ~uint64_t(15);
  };
  // RED_B200_STREAM_WINDOW (bytes, for tests): size of the first window
  uint64_t firstWindow = kFirstWindow;
  if (const char *env = getenv("RED_B200_STREAM_WINDOW")) {
    const uint64_t w = strtoull(env, nullptr, 10);
    if (w >= 4096)
      firstWindow = w & ~uint64_t(15);
  }
  // Last/Full on a program without states that can never be left: nothing can
  // end the walk early, so the whole text is one window
  const bool earlyStyle = (style == RED_B200_STY_INSTANT) || (style == RED_B200_STY_FIRST) ||
                          (style == RED_B200_STY_TANGENT);
  if (!earlyStyle && !pd.hasFinalStates && !getenv("RED_B200_STREAM_WINDOW"))
    firstWindow = (len + 15) & ~uint64_t(15);
  // the most chunks any window can have: 1 KB chunks up to one per lane, more
  // only with 16 KB chunks
  const uint64_t maxChunks = std::max<uint64_t>(lanes, (len + 16383) / 16384) + 1;

  ChunkRec *recs = nullptr, *recs2 = nullptr;
  StreamSummary *dSum = nullptr;
  TailOut *dTail = nullptr;
  CandRec *cands = nullptr;
  BreakRec *breaks = nullptr;
  BreakSummary *dBreak = nullptr;
  ExtraWalk *extras = nullptr;
  unsigned long long *nExtras = nullptr;
  uint8_t *nCand = nullptr, *segIn = nullptr;
  uint32_t *maps = nullptr, *segMaps = nullptr;
  const uint64_t maxSegs = (maxChunks + 31) / 32;
  TRY(scratch_alloc(&recs, maxChunks * sizeof(ChunkRec), stream));
  TRY(scratch_alloc(&recs2, maxChunks * sizeof(ChunkRec), stream));
  TRY(scratch_alloc(&dSum, sizeof(StreamSummary), stream));
  TRY(scratch_alloc(&dTail, sizeof(TailOut), stream));
  TRY(scratch_alloc(&cands, maxChunks * kCand * sizeof(CandRec), stream));
  TRY(scratch_alloc(&nCand, maxChunks, stream));
  TRY(scratch_alloc(&maps, maxChunks * 4, stream));
  TRY(scratch_alloc(&segMaps, maxSegs * 4, stream));
  TRY(scratch_alloc(&segIn, maxSegs, stream));
  TRY(scratch_alloc(&breaks, maxChunks * sizeof(BreakRec), stream));
  TRY(scratch_alloc(&extras, maxChunks * (kCand - 1) * sizeof(ExtraWalk), stream));
  TRY(scratch_alloc(&nExtras, sizeof(unsigned long long), stream));
  TRY(scratch_alloc(&dBreak, sizeof(BreakSummary), stream));
  auto release = [&]() {
    scratch_free(recs, stream);
    scratch_free(recs2, stream);
    scratch_free(dSum, stream);
    scratch_free(dTail, stream);
    scratch_free(cands, stream);
    scratch_free(nCand, stream);
    scratch_free(maps, stream);
    scratch_free(segMaps, stream);
    scratch_free(segIn, stream);
    scratch_free(breaks, stream);
    scratch_free(extras, stream);
    scratch_free(nExtras, stream);
    scratch_free(dBreak, stream);
  };
  struct Releaser {
    decltype(release) &fn;
    ~Releaser() { fn(); }
  } releaser{release};

  auto resultOf = [&](uint32_t v, int32_t &r) -> int {
    uint32_t idx = v / pd.vPerState;
    TRY(cudaMemcpyAsync(&r, pd.results + idx, 4, cudaMemcpyDeviceToHost, stream));
    TRY(cudaStreamSynchronize(stream));
    return RED_B200_OK;
  };

  const uint64_t none = ~0ull;
  StreamSummary init{};
  init.firstBad = none;
  init.firstAccChunk = none;

  /* One window [base, base + bytes) from the exact state v0: every chunk from
     its candidate start states, the maps between neighbouring chunks'
     candidates composed on the device, the true candidate of every chunk
     copied to `into`, and the summary of the window.  track: also record the
     exits from the initial state.  A fixed number of launches and one read
     back; chunks that had no candidates are resolved by one more launch. */
  auto runWindow = [&](uint64_t base, uint64_t bytes, uint32_t v0, uint64_t chunk, bool track,
                       bool stopEarly, ChunkRec *into, StreamSummary &sum, bool &decided) -> int {
    const uint64_t nChunks = (bytes + chunk - 1) / chunk;
    const uint64_t nSegs = (nChunks + 31) / 32;
    const uint8_t *wText = dText + base;  // a multiple of 16 bytes
    // a small window is spread over all SMs with few warps each (a lane
    // walks several times faster when its SM is not full)
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount, (nChunks + 31) / 32));
    const unsigned threads = static_cast<unsigned>(std::min<uint64_t>(
        kMaxThreads, ((nChunks + grid - 1) / grid + 31) / 32 * 32));
    const unsigned cgrid = static_cast<unsigned>((nChunks + 255) / 256);
    const unsigned rgrid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(geo.smCount, (nChunks + 1023) / 1024)));
    auto walk = track ? walkTrack : walkPlain;
    TRY(cudaMemsetAsync(nExtras, 0, sizeof(unsigned long long), stream));
    const bool hotNow = useHot && !track;
    walk<<<grid, threads, smem, stream>>>(pd, wText, bytes, static_cast<uint32_t>(chunk), nChunks, cands, nCand,
                                          v0, kLookback, extras, nExtras, !hotNow);
    if (hotNow) {
      const uint64_t warpsNeeded = (nChunks + 31) / 32;
      const unsigned hgrid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount, warpsNeeded));
      const unsigned hwarps = static_cast<unsigned>(std::min<uint64_t>(hotWarpsMax, (warpsNeeded + hgrid - 1) / hgrid));
      hotWalk<<<hgrid, hwarps * 32, hotFixed, stream>>>(
          pd, *hot, hsp, wText, bytes, static_cast<uint32_t>(chunk), nChunks, cands, nCand);
      ++*launches;
    }
    (track ? extraTrack : extraPlain)<<<grid, threads, smem, stream>>>(pd, wText, bytes, static_cast<uint32_t>(chunk),
                                                                       cands, extras, nExtras);
    ++*launches;
    TRY(cudaMemcpyAsync(dSum, &init, sizeof(init), cudaMemcpyHostToDevice, stream));
    stream_maps_kernel<<<cgrid, 256, 0, stream>>>(cands, nCand, nChunks, maps, segMaps);
    stream_scan_kernel<<<1, 1024, 0, stream>>>(segMaps, nSegs, segIn);
    stream_apply_kernel<<<cgrid, 256, 0, stream>>>(cands, maps, segIn, nChunks, into, dSum);
    *launches += 4;
    {
      // (does nothing unless some chunk had no candidates: then everything
      // from the first of them on, in order)
      if (track)
        stream_repair_kernel<true><<<1, 32, 0, stream>>>(pd, SMEM ? 0u : 1u, wText, bytes,
                                                          static_cast<uint32_t>(chunk), nChunks, cands, nCand, into, dSum);
      else
        stream_repair_kernel<false><<<1, 32, 0, stream>>>(pd, SMEM ? 0u : 1u, wText, bytes,
                                                           static_cast<uint32_t>(chunk), nChunks, cands, nCand, into, dSum);
      ++*launches;
    }
    stream_reduce_kernel<<<rgrid, 256, 0, stream>>>(into, nChunks, nChunks, dSum);
    ++*launches;
    TRY(cudaMemcpyAsync(&sum, dSum, sizeof(sum), cudaMemcpyDeviceToHost, stream));
    TRY(cudaStreamSynchronize(stream));
    // is the walk decided inside this window?
    decided = stopEarly && ((earlyStyle && (sum.firstAccChunk != none)) ||
                            (sum.reduced && (sum.finalV >= pd.absV)));
    return RED_B200_OK;
  };

  struct Win {
    uint64_t base, bytes, chunk;
    uint32_t v0;
  };
  std::vector<Win> wins;  // the windows walked, in text order; recs holds the last one's records

  /* Outcome::start_: the last position in front of `end` at which the initial
     state was left (Matcher.h:436-439).  Nobody tracked it - that costs three
     instructions in every step - so it is searched for now, backwards: the
     tail of the last window in growing pieces (256 KB, x8, each from the exact
     state phase 1 recorded at a chunk boundary) with tracking on, then the
     windows in front of it.  With DFAs of the ".*pattern" kind the first
     piece has it. */
  auto lastLeaveBefore = [&](uint64_t end, uint64_t &found) -> int {
    found = none;
    int rc = RED_B200_OK;
    const Win &L = wins.back();
    uint64_t hiChunk = (end - L.base + L.chunk - 1) / L.chunk;  // chunks [0, hiChunk) of L reach `end`
    uint64_t piece = 256u << 10;
    while ((hiChunk > 0) && (found == none)) {
      const uint64_t nCh = std::max<uint64_t>(1, piece / L.chunk);
      const uint64_t loChunk = (hiChunk > nCh) ? hiChunk - nCh : 0;
      uint32_t vStart = L.v0;
      if (loChunk > 0) {
        ChunkRec r;
        TRY(cudaMemcpyAsync(&r, recs + loChunk, sizeof(r), cudaMemcpyDeviceToHost, stream));
        TRY(cudaStreamSynchronize(stream));
        vStart = r.startV;
      }
      const uint64_t rBase = L.base + loChunk * L.chunk;
      const uint64_t rBytes = std::min(end, L.base + hiChunk * L.chunk) - rBase;
      StreamSummary s2;
      bool dummy = false;
      const uint64_t ch = chunkFor(rBytes);
      if ((rc = runWindow(rBase, rBytes, vStart, ch, true, false, recs2, s2, dummy)) != RED_B200_OK)
        return rc;
      if (s2.lastLeaveP1 != 0)
        found = rBase + (s2.lastLeaveP1 - 1) * ch + s2.lastLeaveOff;
      hiChunk = loChunk;
      piece *= 8;
    }
    for (size_t w = wins.size() - 1; (w-- > 0) && (found == none);) {
      StreamSummary s2;
      bool dummy = false;
      if ((rc = runWindow(wins[w].base, wins[w].bytes, wins[w].v0, wins[w].chunk, true, false, recs2, s2,
                          dummy)) != RED_B200_OK)
        return rc;
      if (s2.lastLeaveP1 != 0)
        found = wins[w].base + (s2.lastLeaveP1 - 1) * wins[w].chunk + s2.lastLeaveOff;
    }
    return RED_B200_OK;
  };

  out->result = 0;
  out->start = 0;
  out->end = 0;
  // carried across windows
  uint32_t v0 = pd.initV;
  uint64_t lastAccEnd = 0;       // one past the last accepting position so far
  uint32_t lastAccV = kNone;
  uint64_t walkedEnd = 0;        // end of the confirmed text

  int rc = RED_B200_OK;
  bool decided = false;
  uint64_t window = firstWindow;
  for (uint64_t base = 0; (base < len) && !decided; base += window, window *= kWindowGrowth) {
    const uint64_t bytes = std::min(window, len - base);
    const uint64_t chunk = chunkFor(bytes);
    wins.push_back(Win{base, bytes, chunk, v0});
    StreamSummary sum;
    if ((rc = runWindow(base, bytes, v0, chunk, false, true, recs, sum, decided)) != RED_B200_OK)
      return rc;

    if (earlyStyle && (sum.firstAccChunk != none)) {
      // careful walk from the chunk with the first accepting position
      const uint64_t c = sum.firstAccChunk;
      const uint64_t from = base + c * chunk;
      const uint64_t until = std::min(len, from + chunk);
      tail<<<1, kMaxThreads, smem, stream>>>(pd, dText, len, from, until, sum.firstAccStartV, style, dTail);
      ++*launches;
      TailOut t;
      TRY(cudaMemcpyAsync(&t, dTail, sizeof(t), cudaMemcpyDeviceToHost, stream));
      TRY(cudaStreamSynchronize(stream));
      if (t.accepted) {
        out->result = t.result;
        out->end = t.end;
        uint64_t leave = t.lastLeave;
        if (t.open) {
          /* Tangent / First: the run of accepting positions outlasts its first
             chunk.  Where it ends is found in parallel: every chunk behind it,
             from its true start state, reports the first position that breaks
             the run - in this window, then in the windows that follow. */
          uint32_t accV = t.accV;
          uint64_t wBase = base, wBytes = bytes, wChunk = chunk, firstChunk = c + 1;
          uint64_t nextBase = base + window, nextWindow = window * kWindowGrowth;
          uint32_t wEndV = sum.finalV;
          bool found = false;
          while (!found) {
            const uint64_t nCh = (wBytes + wChunk - 1) / wChunk;
            if (firstChunk < nCh) {
              const unsigned bgrid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount, (nCh - firstChunk + 31) / 32));
              const unsigned bthreads = static_cast<unsigned>(std::min<uint64_t>(
                  kMaxThreads, ((nCh - firstChunk + bgrid - 1) / bgrid + 31) / 32 * 32));
              brk<<<bgrid, bthreads, smem, stream>>>(pd, dText + wBase, wBytes, static_cast<uint32_t>(wChunk), firstChunk,
                                                     nCh, recs, style, t.result, breaks);
              stream_break_reduce_kernel<<<1, 1024, 0, stream>>>(breaks, firstChunk, nCh, dBreak);
              *launches += 2;
              BreakSummary bs;
              TRY(cudaMemcpyAsync(&bs, dBreak, sizeof(bs), cudaMemcpyDeviceToHost, stream));
              TRY(cudaStreamSynchronize(stream));
              if (bs.leaveChunk != none)
                leave = wBase + bs.leaveChunk * wChunk + bs.leaveOff;
              if (bs.chunk != none) {
                accV = bs.accV;
                out->end = (bs.kind == 2) ? len : wBase + bs.chunk * wChunk + bs.off;
                found = true;
                break;
              }
              accV = bs.endV;
            }
            // the run goes through the window: the next one, from the exact state
            if (nextBase >= len) {
              out->end = len;
              break;
            }
            wBase = nextBase;
            wBytes = std::min(nextWindow, len - wBase);
            wChunk = chunkFor(wBytes);
            StreamSummary s2;
            bool dummy = false;
            if ((rc = runWindow(wBase, wBytes, wEndV, wChunk, false, false, recs, s2, dummy)) != RED_B200_OK)
              return rc;
            wins.push_back(Win{wBase, wBytes, wChunk, wEndV});
            wEndV = s2.finalV;
            firstChunk = 0;
            nextBase = wBase + nextWindow;
            nextWindow *= kWindowGrowth;
          }
          if (accV != t.accV) {
            int32_t r2 = t.result;
            if ((rc = resultOf(accV, r2)) != RED_B200_OK)
              return rc;
            out->result = r2;  // (Tangent reports the last accepting state's result)
          }
        }
        // the initial state was last left either inside the walk from `from`
        // or somewhere in front of it
        if ((leave == none) && (from > 0)) {
          // lastLeaveBefore looks at the window `from` lies in - whose records
          // it expects in recs - and at those in front of it
          if (wins.back().base > from) {
            while (wins.back().base > from)
              wins.pop_back();
            StreamSummary s2;
            bool dummy = false;
            const Win &w = wins.back();
            if ((rc = runWindow(w.base, w.bytes, w.v0, w.chunk, false, false, recs, s2, dummy)) != RED_B200_OK)
              return rc;
          }
          if ((rc = lastLeaveBefore(from, leave)) != RED_B200_OK)
            return rc;
        }
        out->start = (leave == none) ? 0 : leave;
      }
      return RED_B200_OK;
    }

    // fold the window (its confirmed prefix, which is all of it unless the
    // walk was decided) into the running outcome
    if (sum.lastAccP1 != 0) {
      lastAccEnd = base + (sum.lastAccP1 - 1) * chunk + sum.lastAccEnd;
      lastAccV = sum.lastAccV;
    }
    walkedEnd = std::min(len, base + sum.reduced * chunk);
    v0 = sum.finalV;
    if (v0 >= pd.absV)
      decided = true;
  }

  // Instant/First/Tangent come here only without a match.  v0 is the final
  // state, or a state that can never be left.
  const bool absorbing = (v0 >= pd.absV) && (v0 < pd.pureV);
  int32_t r = 0;
  uint32_t resultV = kNone;
  if (style == RED_B200_STY_FULL) {
    // the final state decides (a pure dead end carries result 0)
    if ((v0 >= pd.acceptV) && (v0 < pd.pureV))
      resultV = v0;
  }
  else if (style == RED_B200_STY_LAST)
    resultV = lastAccV;
  if (resultV != kNone) {
    if ((rc = resultOf(resultV, r)) != RED_B200_OK)
      return rc;
    uint64_t leave = none;
    if ((r > 0) && (walkedEnd > 0))
      if ((rc = lastLeaveBefore(walkedEnd, leave)) != RED_B200_OK)
        return rc;
    out->result = r;
    out->start = (leave == none) ? 0 : leave;
    out->end = ((style == RED_B200_STY_FULL) || absorbing) ? len : lastAccEnd;
  }
#undef TRY
  return RED_B200_OK;
}

} // anonymous

int match_stream(const ProgramDev &pd, bool smemTable, const DeviceImage &, const HotDev *hot,
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
      return cls8 ? run_stream<true, 1, uint8_t>(pd, hot, geo, dText, len, style, out, stream, launches, why)
                  : run_stream<true, 1, uint16_t>(pd, hot, geo, dText, len, style, out, stream, launches, why);
    if (pd.unitShift == 1)
      return cls8 ? run_stream<true, 2, uint8_t>(pd, hot, geo, dText, len, style, out, stream, launches, why)
                  : run_stream<true, 2, uint16_t>(pd, hot, geo, dText, len, style, out, stream, launches, why);
    if (pd.unitShift == 8)
      return run_stream<true, 9, uint16_t>(pd, hot, geo, dText, len, style, out, stream, launches, why);
    why = "unsupported table layout";
    return RED_B200_EINTERNAL;
  }
  return cls8 ? run_stream<false, 0, uint8_t>(pd, hot, geo, dText, len, style, out, stream, launches, why)
              : run_stream<false, 0, uint16_t>(pd, hot, geo, dText, len, style, out, stream, launches, why);
}

} // namespace red_b200
