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
     2. a window is cut into chunks, one lane per chunk (2-16 KB each);
     3. SPECULATE the state at the start of chunk i: run the DFA from the
        initial state over a lookback window in front of the chunk.  DFAs of
        the ".*pattern" kind forget their past quickly, so the guess is almost
        always the true state;
     4. walk every chunk from its guess, in parallel, with the batch kernels'
        chunk engine (chunk_engine.cuh) in style Last with start tracking: the
        record of a chunk is its end state, its last accepting position with
        the state there, and the last position at which the initial state was
        left;
     5. STITCH: chunk i is confirmed when its start state equals the end state
        of chunk i-1.  The records in front of the first unconfirmed chunk are
        reduced in text order; if they already decide the walk, it is over.
        Otherwise the unconfirmed chunks get the correct start state and are
        walked again - all of them in parallel - until a round changes
        nothing.  A chunk whose guess was wrong usually still ENDS in the
        right state (the DFA converges inside the chunk), so a wrong guess
        does not cascade and the loop ends after one or two rounds; in the
        worst case (a DFA that never forgets, e.g. a parity counter) it
        degrades to one chunk per round and stays exact;
     6. the styles that stop at the first match get one short careful walk
        from the start of the chunk that holds it.

   Every answer is exact: nothing is derived from an unconfirmed chunk.
 */

#include "stream.h"
#include "chunk_engine.cuh"

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

/* text/len: the window; v0: the (exact) state at its start */
template <bool SMEM, int K, class ClsT, bool TRACK>
__global__ void __launch_bounds__(kMaxThreads, 1)
stream_walk_kernel(const ProgramDev pd, const uint8_t *text, uint64_t len,
                   uint32_t chunkBytes, uint64_t nChunks, ChunkRec *recs,
                   int firstRound, uint32_t v0) {
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
  for (uint64_t r = 0; r < rounds; ++r) {
    const uint64_t c = first + r * stride;
    ChunkRec rec{};
    bool mine = c < nChunks;
    if (mine && !firstRound) {
      rec = recs[c];
      mine = rec.dirty != 0;
    }
    const uint64_t begin = mine ? c * chunkBytes : 0;
    const uint32_t n = mine ? static_cast<uint32_t>(min(static_cast<uint64_t>(chunkBytes), len - begin)) : 0;
    if (firstRound) {
      // speculate: the state after the lookback window, entered from the
      // initial state (chunk 0 starts in the known state v0)
      Walk<STY_FULL, false> sp;
      sp.begin(pd);
      const uint32_t lb = (mine && (c > 0)) ? kLookback : 0;  // chunkBytes >= kLookback
      walk_bytes<STY_FULL, false, Eng>(pd, fast, tab, cls, text + begin - lb, lb, sp);
      rec.startV = (c == 0) ? v0 : sp.v;
    }
    // TRACK: also the last exit from the initial state (three more
    // instructions per step: only the lazy start search asks for it)
    Walk<STY_LAST, TRACK> wk;
    wk.begin(pd);
    wk.v = rec.startV;
    wk.mstart = kNoPos;
    walk_bytes<STY_LAST, TRACK, Eng>(pd, fast, tab, cls, text + begin, n, wk);
    if (mine) {
      rec.endV = wk.v;
      rec.accV = wk.accV;
      rec.accEnd = wk.absorbed ? n : wk.mend;
      rec.lastLeaveOff = TRACK ? wk.mstart : kNoPos;
      rec.dirty = 0;
      recs[c] = rec;
    }
  }
}

/* confirm chunk i against the end state of chunk i-1; hand the right start
   state to the ones that fail, count them and find the first */
__global__ void stream_verify_kernel(ChunkRec *recs, uint64_t nChunks,
                                     uint32_t v0, StreamSummary *sum) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  unsigned bad = 0;
  unsigned long long firstBad = ~0ull;
  for (uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       c < nChunks; c += stride) {
    const uint32_t want = (c == 0) ? v0 : recs[c - 1].endV;
    if (recs[c].startV != want) {
      recs[c].startV = want;
      recs[c].dirty = 1;
      ++bad;
      firstBad = min(firstBad, static_cast<unsigned long long>(c));
    }
  }
  __syncwarp();
  if (__any_sync(0xffffffffu, bad != 0)) {
    bad = __reduce_add_sync(0xffffffffu, bad);
    for (int d = 16; d > 0; d >>= 1)
      firstBad = min(firstBad, __shfl_xor_sync(0xffffffffu, firstBad, d));
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sum->bad, static_cast<unsigned long long>(bad));
      atomicMin(&sum->firstBad, firstBad);
    }
  }
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

  const size_t smem = kLanesSmemExtra + (SMEM ? pd.tableBytes : pd.hotBytes);
  auto walkPlain = stream_walk_kernel<SMEM, K, ClsT, false>;
  auto walkTrack = stream_walk_kernel<SMEM, K, ClsT, true>;
  auto tail = stream_tail_kernel<SMEM, K, ClsT>;
  TRY(cudaFuncSetAttribute(walkPlain, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(walkTrack, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  TRY(cudaFuncSetAttribute(tail, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));

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
  TRY(scratch_alloc(&recs, maxChunks * sizeof(ChunkRec), stream));
  TRY(scratch_alloc(&recs2, maxChunks * sizeof(ChunkRec), stream));
  TRY(scratch_alloc(&dSum, sizeof(StreamSummary), stream));
  TRY(scratch_alloc(&dTail, sizeof(TailOut), stream));
  auto release = [&]() {
    scratch_free(recs, stream);
    scratch_free(recs2, stream);
    scratch_free(dSum, stream);
    scratch_free(dTail, stream);
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

  /* One window [base, base + bytes) from the exact state v0: walk, confirm,
     re-walk until nothing changes (or, with stopEarly, until the confirmed
     prefix decides the walk).  track: also record the exits from the initial
     state.  The records stay in `into`. */
  auto runWindow = [&](uint64_t base, uint64_t bytes, uint32_t v0, uint64_t chunk, bool track,
                       bool stopEarly, ChunkRec *into, StreamSummary &sum, bool &decided) -> int {
    const uint64_t nChunks = (bytes + chunk - 1) / chunk;
    const uint8_t *wText = dText + base;  // a multiple of 16 bytes
    // a small window is spread over all SMs with few warps each (a lane
    // walks several times faster when its SM is not full)
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount, (nChunks + 31) / 32));
    const unsigned threads = static_cast<unsigned>(std::min<uint64_t>(
        kMaxThreads, ((nChunks + grid - 1) / grid + 31) / 32 * 32));
    const unsigned vgrid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount * 4, (nChunks + 255) / 256));
    const unsigned rgrid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(geo.smCount, (nChunks + 1023) / 1024)));
    auto walk = track ? walkTrack : walkPlain;
    walk<<<grid, threads, smem, stream>>>(pd, wText, bytes, static_cast<uint32_t>(chunk), nChunks, into, 1, v0);
    ++*launches;
    decided = false;
    for (uint64_t round = 0;; ++round) {
      TRY(cudaMemcpyAsync(dSum, &init, sizeof(init), cudaMemcpyHostToDevice, stream));
      stream_verify_kernel<<<vgrid, 256, 0, stream>>>(into, nChunks, v0, dSum);
      stream_reduce_kernel<<<rgrid, 256, 0, stream>>>(into, nChunks, none, dSum);
      *launches += 2;
      TRY(cudaMemcpyAsync(&sum, dSum, sizeof(sum), cudaMemcpyDeviceToHost, stream));
      TRY(cudaStreamSynchronize(stream));
      // is the walk decided inside the confirmed prefix?
      if (stopEarly && ((earlyStyle && (sum.firstAccChunk != none)) ||
                        (sum.reduced && (sum.finalV >= pd.absV)))) {
        decided = true;
        break;
      }
      if (sum.bad == 0)
        break;
      if (round > nChunks) {
        why = "stream stitching did not converge";
        return RED_B200_EINTERNAL;
      }
      walk<<<grid, threads, smem, stream>>>(pd, wText, bytes, static_cast<uint32_t>(chunk), nChunks, into, 0, v0);
      ++*launches;
    }
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
      tail<<<1, kMaxThreads, smem, stream>>>(pd, dText, len, from, sum.firstAccStartV, style, dTail);
      ++*launches;
      TailOut t;
      TRY(cudaMemcpyAsync(&t, dTail, sizeof(t), cudaMemcpyDeviceToHost, stream));
      TRY(cudaStreamSynchronize(stream));
      if (t.accepted) {
        out->result = t.result;
        out->end = t.end;
        // the initial state was last left either inside the tail walk or
        // somewhere in front of it
        uint64_t leave = t.lastLeave;
        if ((leave == none) && (from > 0))
          if ((rc = lastLeaveBefore(from, leave)) != RED_B200_OK)
            return rc;
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
