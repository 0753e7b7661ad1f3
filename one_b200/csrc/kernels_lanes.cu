/* kernels_lanes.cu - the production batch kernels for check/match.

   Two schedules share one chunk engine (process_chunk):
     batch_tile_kernel   (default)  a warp takes 32 consecutive lines, one per
                                    lane, all lanes in step, outcomes stored
                                    together; next tile prefetched
     batch_lanes_kernel  (RED_B200_VARIANT=3)  persistent lanes, described next
   They measure within a few percent of each other on B200 (profiles/); the
   tile schedule is simpler and slightly ahead, so it is the default.

   Persistent lanes.

   Why: with one line per lane per tile (kernels.cu, kept as the A/B baseline)
   ncu showed 11.8 of 32 lanes active per instruction - lines differ in length,
   matched lines stop early, and the head/tail chunks of a line ran as separate
   divergent passes - while the shared-memory (LSU) pipe that serves the two
   table loads per byte was already 70% busy.  Every wavefront and every issue
   slot was mostly empty lanes.

   Here a warp claims a block of 32*T consecutive lines from a global counter;
   lane l owns lines base+l, base+l+32, ... and streams through them on its
   own: when its line ends it writes the outcome and moves to its next line in
   the same loop iteration, so all lanes stay busy until the block drains.
   One loop iteration = one aligned 16-byte chunk per lane:

     * a per-lane 16-bit mask says which bytes of the chunk belong to the line
       (head and tail chunks are partial), so there is one code path;
     * the chunk is walked without decisions; per 4-byte quarter only the
       maximum state value is kept.  If no quarter reached `limit` (the common
       case) the chunk is done;  otherwise only the flagged quarters are
       re-walked with the reference's per-byte decisions (walk.cuh);
     * the next line's metadata and first chunk, and the current line's next
       chunk, are requested one step ahead.

   Semantics per line are those of matchCore/checkCore (Matcher.h:363-495).
 */

#include "chunk_engine.cuh"

#include <algorithm>
#include <type_traits>

namespace red_b200 {

namespace {

using namespace dev;

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT, bool LEADER>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_lanes_kernel(const ProgramDev pd, const BatchArgs args,
                   unsigned long long *blockCounter, const uint32_t T) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  using Eng = Engine<SMEM, K, ClsT>;
  unsigned char *smem = Eng::stage(pd, smemRaw);
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const typename Eng::Tab tab = Eng::table(pd, smem);
  const typename Eng::Fast fast = Eng::fast(pd, smem);

  const uint32_t lane = threadIdx.x & 31;
  const uint64_t perBlock = 32ull * T;
  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint32_t leadLen = LEADER ? pd.leaderLen : 0;

  for (;;) {
    unsigned long long blk = 0;
    if (lane == 0)
      blk = atomicAdd(blockCounter, 1ull);
    blk = __shfl_sync(kFull, blk, 0);
    const uint64_t base = blk * perBlock;
    if (base >= n)
      break;
    // this block's lines are base + [0, count); lane l owns offsets l, l+32, ...
    const uint32_t count = static_cast<uint32_t>(min(n - base, perBlock));
    const uint64_t *starts = args.starts + base;
    const uint32_t *lens = args.lens + base;

    // off: offset of the current line (C).  B = off + 32 is the next line, held
    // with its first chunk; A = off + 64 is the one after, metadata only.
    uint32_t off = lane;
    const uint8_t *aPtr = nullptr, *bPtr = nullptr;
    uint32_t aLen = 0, bLen = 0;
    uint4 bFirst = make_uint4(0, 0, 0, 0);
    const uint4 *cBlk = nullptr;
    uint32_t cSkip = 0, cTotal = 0, cK = 0;
    uint4 cur = make_uint4(0, 0, 0, 0), nxt = make_uint4(0, 0, 0, 0);
    bool cNoResult = false;  // check(): the leader did not match
    Walk<STYLE, WANT_START> wk;
    wk.begin(pd);

    auto meta = [&](uint32_t o, const uint8_t *&ptr, uint32_t &len) {
      if (o < count) {
        ptr = args.text + __ldg(starts + o);
        len = __ldg(lens + o);
      }
    };
    auto firstChunk = [&](uint32_t o, const uint8_t *ptr, uint32_t len) {
      uint4 q = make_uint4(0, 0, 0, 0);
      if ((o < count) && (len > leadLen)) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(ptr + leadLen);
        q = __ldg(reinterpret_cast<const uint4 *>(a & ~uintptr_t(15)));
      }
      return q;
    };
    // B becomes the current line (its offset is already in `off`); A becomes
    // B; metadata for the new A is requested
    auto advance = [&]() {
      wk.begin(pd);  // also for a lane that has run out of lines: a stale
                     // accepting state would keep flagging its (empty) chunks
      if (off < count) {
        const uint8_t *p = bPtr;
        uint32_t len = bLen;
        if (LEADER) {
          // check() with the leader (Matcher.h:369-376): compare the required
          // prefix, then start after it from the state it leads to
          bool viable = true;
          for (uint32_t ii = 0; ii < leadLen; ++ii)
            if ((ii >= len) || (pd.leader[ii] != pd.equiv[__ldg(p + ii)])) {
              viable = false;
              break;
            }
          p += leadLen;
          len = viable ? len - leadLen : 0;
          wk.v = pd.leaderV;
          cNoResult = !viable;  // result 0, nothing to walk
        }
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        cSkip = static_cast<uint32_t>(a & 15);
        cBlk = reinterpret_cast<const uint4 *>(a - cSkip);
        cTotal = cSkip + len;
        cK = 0;
        cur = bFirst;
      }
      bPtr = aPtr;
      bLen = aLen;
      bFirst = firstChunk(off + 32, bPtr, bLen);
      meta(off + 64, aPtr, aLen);
    };

    // prime the pipeline
    meta(off, bPtr, bLen);
    bFirst = firstChunk(off, bPtr, bLen);
    meta(off + 32, aPtr, aLen);
    advance();

    while (__any_sync(kFull, off < count)) {
      const bool cValid = off < count;
      uint32_t lo = 0, hi = 0;
      bool lastChunk = true;
      if (cValid) {
        const uint32_t rem = cTotal - cK * 16;
        lo = (cK == 0) ? cSkip : 0;
        hi = min(16u, rem);
        lastChunk = rem <= 16;
        if (!lastChunk)
          nxt = __ldg(cBlk + cK + 1);
      }
      const uint32_t mask = (0xffffu >> (16 - hi)) & (0xffffu << lo);
      const uint32_t pos0 = cK * 16 - cSkip;  // line position of byte 0 (mod 2^32)
      bool stopped = false;

      stopped = process_chunk<STYLE, WANT_START>(pd, fast, tab, cls, cur, mask, pos0, wk, cValid);

      if (cValid) {
        if (stopped || lastChunk) {
          int32_t result = 0;
          uint32_t start = 0, end = 0;
          if (!(LEADER && cNoResult))
            wk.finish(pd, cTotal - cSkip, result, start, end);
          const uint64_t line = base + off;
          args.results[line] = result;
          if (WANT_START)
            args.matchStarts[line] = start;
          if (args.matchEnds)
            args.matchEnds[line] = end;
          off += 32;
          advance();
        }
        else {
          ++cK;
          cur = nxt;
        }
      }
    }
  }
}

/* batch_tile_kernel: the same chunk machinery, different scheduling.  A warp
   claims 32 consecutive lines, one per lane; all lanes walk in step until the
   last one is done; then all outcomes are finished and stored together
   (coalesced), and the warp moves to its next tile, whose metadata and first
   chunks were requested while the current tile was being walked.

   Against the persistent-lane schedule: lanes whose line is over idle until
   the longest line of the tile ends (slot use ~ avg/max chunks), but nothing
   per line is executed inside the chunk loop at 4-of-32 lane efficiency, the
   index loads and result stores are coalesced, and lanes at the same depth of
   similar lines hit the same table words more often (fewer bank conflicts). */
template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT, bool LEADER>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_tile_kernel(const ProgramDev pd, const BatchArgs args,
                  unsigned long long *blockCounter) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  using Eng = Engine<SMEM, K, ClsT>;
  unsigned char *smem = Eng::stage(pd, smemRaw);
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const typename Eng::Tab tab = Eng::table(pd, smem);
  const typename Eng::Fast fast = Eng::fast(pd, smem);

  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint32_t leadLen = LEADER ? pd.leaderLen : 0;

  auto claim = [&]() -> uint64_t {
    unsigned long long t = 0;
    if (lane == 0)
      t = atomicAdd(blockCounter, 1ull);
    return __shfl_sync(kFull, t, 0) * 32ull + lane;
  };
  // next tile: its line for this lane, with the first chunk
  uint64_t nLine = claim();
  const uint8_t *nPtr = nullptr;
  uint32_t nLen = 0;
  uint4 nFirst = make_uint4(0, 0, 0, 0);
  auto fetchMeta = [&]() {
    if (nLine < n) {
      nPtr = args.text + __ldg(args.starts + nLine);
      nLen = __ldg(args.lens + nLine);
    }
  };
  auto fetchFirst = [&]() {
    if ((nLine < n) && (nLen > leadLen)) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(nPtr + leadLen);
      nFirst = __ldg(reinterpret_cast<const uint4 *>(a & ~uintptr_t(15)));
    }
  };
  fetchMeta();
  fetchFirst();

  while (__any_sync(kFull, nLine < n)) {
    // the prefetched tile becomes the current one
    const uint64_t line = nLine;
    bool active = line < n;
    const uint8_t *p = nPtr;
    uint32_t len = active ? nLen : 0;
    uint4 cur = nFirst, nxt = nFirst;
    Walk<STYLE, WANT_START> wk;
    wk.begin(pd);
    bool noResult = false;
    if (LEADER && active) {
      // check() with the leader (Matcher.h:369-376)
      bool viable = true;
      for (uint32_t ii = 0; ii < leadLen; ++ii)
        if ((ii >= len) || (pd.leader[ii] != pd.equiv[__ldg(p + ii)])) {
          viable = false;
          break;
        }
      p += leadLen;
      len = viable ? len - leadLen : 0;
      wk.v = pd.leaderV;
      noResult = !viable;
    }
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t skip = active ? static_cast<uint32_t>(a & 15) : 0;
    const uint4 *blk = reinterpret_cast<const uint4 *>(a - skip);
    const uint32_t total = skip + len;
    const uint32_t lineLen = len;

    // claim the tile after this one and ask for its index entries now; its
    // first chunks are requested one iteration later, when the entries are in
    nLine = claim();
    fetchMeta();
    bool wantFirst = true;

    uint32_t k = 0;
    while (__any_sync(kFull, active)) {
      uint32_t lo = 0, hi = 0;
      bool lastChunk = true;
      if (active) {
        const uint32_t rem = total - k * 16;
        lo = (k == 0) ? skip : 0;
        hi = min(16u, rem);
        lastChunk = rem <= 16;
        if (!lastChunk)
          nxt = __ldg(blk + k + 1);
      }
      if (wantFirst && (k == 1)) {
        fetchFirst();
        wantFirst = false;
      }
      const uint32_t mask = (0xffffu >> (16 - hi)) & (0xffffu << lo);
      const uint32_t pos0 = k * 16 - skip;
      const bool stopped =
          process_chunk<STYLE, WANT_START>(pd, fast, tab, cls, cur, mask, pos0, wk, active);
      if (stopped || lastChunk)
        active = false;
      ++k;
      cur = nxt;
    }
    if (wantFirst)
      fetchFirst();

    if (line < n) {
      int32_t result = 0;
      uint32_t start = 0, end = 0;
      if (!(LEADER && noResult))
        wk.finish(pd, lineLen, result, start, end);
      args.results[line] = result;
      if (WANT_START)
        args.matchStarts[line] = start;
      if (args.matchEnds)
        args.matchEnds[line] = end;
    }
  }
}

/* batch_tile32_kernel: the tile schedule with 32-byte text loads
   (LDG.E.256, sm_100).  A lane's text load costs one L1 data-pipe wavefront
   whatever its width (every lane is in a different 128-byte line), and those
   wavefronts are a quarter of the LSU pipe's work in the 16-byte version
   (ncu: l1tex__data_pipe_lsu_wavefronts_mem_lgds vs _mem_shared); loading 32
   bytes at a time halves them.  Each 32-byte block is walked as two 16-byte
   chunks by the same engine. */
struct Text32 {
  uint4 lo, hi;
};
__device__ __forceinline__ Text32 ldg256(const void *p) {
  Text32 t;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(t.lo.x), "=r"(t.lo.y), "=r"(t.lo.z), "=r"(t.lo.w),
                 "=r"(t.hi.x), "=r"(t.hi.y), "=r"(t.hi.z), "=r"(t.hi.w)
               : "l"(p));
  return t;
}

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT, bool LEADER>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_tile32_kernel(const ProgramDev pd, const BatchArgs args,
                    unsigned long long *blockCounter) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  using Eng = Engine<SMEM, K, ClsT>;
  unsigned char *smem = Eng::stage(pd, smemRaw);
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const typename Eng::Tab tab = Eng::table(pd, smem);
  const typename Eng::Fast fast = Eng::fast(pd, smem);

  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint32_t leadLen = LEADER ? pd.leaderLen : 0;

  auto claim = [&]() -> uint64_t {
    unsigned long long t = 0;
    if (lane == 0)
      t = atomicAdd(blockCounter, 1ull);
    return __shfl_sync(kFull, t, 0) * 32ull + lane;
  };
  uint64_t nLine = claim();
  const uint8_t *nPtr = nullptr;
  uint32_t nLen = 0;
  Text32 nFirst;
  nFirst.lo = nFirst.hi = make_uint4(0, 0, 0, 0);
  auto fetchMeta = [&]() {
    if (nLine < n) {
      nPtr = args.text + __ldg(args.starts + nLine);
      nLen = __ldg(args.lens + nLine);
    }
  };
  auto fetchFirst = [&]() {
    if ((nLine < n) && (nLen > leadLen)) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(nPtr + leadLen);
      nFirst = ldg256(reinterpret_cast<const void *>(a & ~uintptr_t(31)));
    }
  };
  fetchMeta();
  fetchFirst();

  while (__any_sync(kFull, nLine < n)) {
    const uint64_t line = nLine;
    bool active = line < n;
    const uint8_t *p = nPtr;
    uint32_t len = active ? nLen : 0;
    Text32 cur = nFirst, nxt = nFirst;
    Walk<STYLE, WANT_START> wk;
    wk.begin(pd);
    bool noResult = false;
    if (LEADER && active) {
      // check() with the leader (Matcher.h:369-376)
      bool viable = true;
      for (uint32_t ii = 0; ii < leadLen; ++ii)
        if ((ii >= len) || (pd.leader[ii] != pd.equiv[__ldg(p + ii)])) {
          viable = false;
          break;
        }
      p += leadLen;
      len = viable ? len - leadLen : 0;
      wk.v = pd.leaderV;
      noResult = !viable;
    }
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t skip = active ? static_cast<uint32_t>(a & 31) : 0;
    const char *blk = reinterpret_cast<const char *>(a - skip);
    const uint32_t total = skip + len;
    const uint32_t lineLen = len;

    nLine = claim();
    fetchMeta();
    bool wantFirst = true;

    uint32_t k = 0;
    while (__any_sync(kFull, active)) {
      uint32_t mask32 = 0;
      bool lastBlock = true;
      if (active) {
        const uint32_t rem = total - k * 32;
        const uint32_t lo = (k == 0) ? skip : 0;
        const uint32_t hi = min(32u, rem);
        lastBlock = rem <= 32;
        if (!lastBlock)
          nxt = ldg256(blk + (k + 1) * 32);
        mask32 = (0xffffffffu >> (32 - hi)) & (0xffffffffu << lo);
      }
      if (wantFirst && (k == 1)) {
        fetchFirst();
        wantFirst = false;
      }
      const uint32_t pos0 = k * 32 - skip;
      bool stopped = process_chunk<STYLE, WANT_START>(pd, fast, tab, cls, cur.lo, mask32 & 0xffffu,
                                                      pos0, wk, active);
      if (!stopped)
        stopped = process_chunk<STYLE, WANT_START>(pd, fast, tab, cls, cur.hi, mask32 >> 16,
                                                   pos0 + 16, wk, active);
      if (stopped || lastBlock)
        active = false;
      ++k;
      cur = nxt;
    }
    if (wantFirst)
      fetchFirst();

    if (line < n) {
      int32_t result = 0;
      uint32_t start = 0, end = 0;
      if (!(LEADER && noResult))
        wk.finish(pd, lineLen, result, start, end);
      args.results[line] = result;
      if (WANT_START)
        args.matchStarts[line] = start;
      if (args.matchEnds)
        args.matchEnds[line] = end;
    }
  }
}

/* batch_sorted_kernel (default when positions are tracked; RED_B200_VARIANT=4
   forces it): the tile schedule with tiles of SIMILAR
   lines.  In a tile every lane waits for the longest line; with consecutive
   lines of a log one instruction serves 20 of 32 lanes on average (ncu:
   profiles/), and the shared-memory pipe that bounds the kernel is busy with
   mostly-empty wavefronts.  Here a warp claims a GROUP of 32*T consecutive
   lines, buckets them by length in 16-byte chunks (a counting sort in its
   slice of shared memory: two passes of shared-memory atomics, 2 per line),
   and walks the group as T tiles of lines with (nearly) the same chunk count.
   Outcomes go to the line's own slot, so the order of the results is
   untouched.  The next group is claimed and its index entries pulled into L2
   a whole group ahead; it is sorted just before the last tile of the current
   group is walked, so the prefetch chain (index entry -> first chunk) never
   drains. */
constexpr uint32_t kNoIdx = 0xffffu;

// bytes of shared memory the sort needs per CTA of `warps` warps
inline size_t sorted_smem_bytes(int warps, uint32_t T) {
  return static_cast<size_t>(warps) * (32u * T * 2u + 32u * 4u);
}

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT, bool LEADER>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_sorted_kernel(const ProgramDev pd, const BatchArgs args,
                    unsigned long long *groupCounter, const uint32_t T,
                    const uint32_t sortOffset) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  using Eng = Engine<SMEM, K, ClsT>;
  unsigned char *smem = Eng::stage(pd, smemRaw);
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const typename Eng::Tab tab = Eng::table(pd, smem);
  const typename Eng::Fast fast = Eng::fast(pd, smem);

  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t nWarps = blockDim.x >> 5;
  // sort scratch of this warp: bucket counters, then the sorted line offsets
  uint32_t *hist = reinterpret_cast<uint32_t *>(smem + sortOffset) + warp * 32;
  uint16_t *order = reinterpret_cast<uint16_t *>(smem + sortOffset + nWarps * 128) + warp * 32 * T;

  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint32_t leadLen = LEADER ? pd.leaderLen : 0;
  const uint64_t perGroup = 32ull * T;

  auto claimGroup = [&]() -> uint64_t {
    unsigned long long g = 0;
    if (lane == 0)
      g = atomicAdd(groupCounter, 1ull);
    return __shfl_sync(kFull, g, 0) * perGroup;
  };
  // pull a group's index entries into L2 ahead of use
  auto prefetchGroup = [&](uint64_t base) {
    if (base >= n)
      return;
    const uint64_t cnt = min(n - base, perGroup);
    const char *ls = reinterpret_cast<const char *>(args.lens + base);
    const char *ss = reinterpret_cast<const char *>(args.starts + base);
    for (uint64_t o = lane * 128ull; o < cnt * 4; o += 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ls + o));
    for (uint64_t o = lane * 128ull; o < cnt * 8; o += 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ss + o));
  };
  auto bucketOf = [&](uint32_t len) -> uint32_t { return min(31u, (len + 15u) >> 4); };
  // counting sort of the group's lines by chunk count -> order[]
  auto prepare = [&](uint64_t base) {
    const uint32_t cnt = static_cast<uint32_t>(min(n - base, perGroup));
    hist[lane] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < cnt; i += 32)
      atomicAdd(&hist[bucketOf(__ldg(args.lens + base + i))], 1u);
    __syncwarp();
    // exclusive scan of the 32 counters
    const uint32_t c = hist[lane];
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(kFull, incl, d);
      if (lane >= d)
        incl += up;
    }
    __syncwarp();
    hist[lane] = incl - c;
    __syncwarp();
    for (uint32_t i = lane; i < cnt; i += 32) {
      const uint32_t pos = atomicAdd(&hist[bucketOf(__ldg(args.lens + base + i))], 1u);
      order[pos] = static_cast<uint16_t>(i);
    }
    for (uint32_t i = cnt + lane; i < perGroup; i += 32)
      order[i] = kNoIdx;
    __syncwarp();
  };

  uint64_t gBase = claimGroup();
  uint64_t gNext = claimGroup();
  uint32_t t = 0;
  bool more = gBase < n;
  if (more)
    prepare(gBase);
  prefetchGroup(gNext);

  // next tile: its line for this lane (~0: none), with the first chunk
  uint64_t nLine = ~0ull;
  const uint8_t *nPtr = nullptr;
  uint32_t nLen = 0;
  uint4 nFirst = make_uint4(0, 0, 0, 0);
  auto pickTile = [&]() {
    const uint32_t idx = order[t * 32 + lane];
    nLine = (idx != kNoIdx) ? gBase + idx : ~0ull;
  };
  auto fetchMeta = [&]() {
    if (nLine != ~0ull) {
      nPtr = args.text + __ldg(args.starts + nLine);
      nLen = __ldg(args.lens + nLine);
    }
  };
  auto fetchFirst = [&]() {
    if ((nLine != ~0ull) && (nLen > leadLen)) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(nPtr + leadLen);
      nFirst = __ldg(reinterpret_cast<const uint4 *>(a & ~uintptr_t(15)));
    }
  };
  if (more) {
    pickTile();
    fetchMeta();
    fetchFirst();
  }

  while (more) {
    // the prefetched tile becomes the current one
    const uint64_t line = nLine;
    bool active = line != ~0ull;
    const uint8_t *p = nPtr;
    uint32_t len = active ? nLen : 0;
    uint4 cur = nFirst, nxt = nFirst;
    Walk<STYLE, WANT_START> wk;
    wk.begin(pd);
    bool noResult = false;
    if (LEADER && active) {
      // check() with the leader (Matcher.h:369-376)
      bool viable = true;
      for (uint32_t ii = 0; ii < leadLen; ++ii)
        if ((ii >= len) || (pd.leader[ii] != pd.equiv[__ldg(p + ii)])) {
          viable = false;
          break;
        }
      p += leadLen;
      len = viable ? len - leadLen : 0;
      wk.v = pd.leaderV;
      noResult = !viable;
    }
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t skip = active ? static_cast<uint32_t>(a & 15) : 0;
    const uint4 *blk = reinterpret_cast<const uint4 *>(a - skip);
    const uint32_t total = skip + len;
    const uint32_t lineLen = len;

    // move on to the tile after this one (sorting the next group when this
    // was the last tile of its group) and ask for its index entries now; its
    // first chunks are requested one iteration later, when the entries are in
    if (++t == T) {
      t = 0;
      gBase = gNext;
      more = gBase < n;
      if (more) {
        prepare(gBase);
        gNext = claimGroup();
        prefetchGroup(gNext);
      }
    }
    nLine = ~0ull;
    if (more) {
      pickTile();
      fetchMeta();
    }
    bool wantFirst = true;

    uint32_t k = 0;
    while (__any_sync(kFull, active)) {
      uint32_t lo = 0, hi = 0;
      bool lastChunk = true;
      if (active) {
        const uint32_t rem = total - k * 16;
        lo = (k == 0) ? skip : 0;
        hi = min(16u, rem);
        lastChunk = rem <= 16;
        if (!lastChunk)
          nxt = __ldg(blk + k + 1);
      }
      if (wantFirst && (k == 1)) {
        fetchFirst();
        wantFirst = false;
      }
      const uint32_t mask = (0xffffu >> (16 - hi)) & (0xffffu << lo);
      const uint32_t pos0 = k * 16 - skip;
      const bool stopped =
          process_chunk<STYLE, WANT_START>(pd, fast, tab, cls, cur, mask, pos0, wk, active);
      if (stopped || lastChunk)
        active = false;
      ++k;
      cur = nxt;
    }
    if (wantFirst)
      fetchFirst();

    if (line != ~0ull) {
      int32_t result = 0;
      uint32_t start = 0, end = 0;
      if (!(LEADER && noResult))
        wk.finish(pd, lineLen, result, start, end);
      args.results[line] = result;
      if (WANT_START)
        args.matchStarts[line] = start;
      if (args.matchEnds)
        args.matchEnds[line] = end;
    }
  }
}

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT>
cudaError_t launch_lanes(const ProgramDev &pd, const BatchArgs &args,
                         const LaunchGeometry &geo, const BatchTuning &tune,
                         unsigned long long *counter, cudaStream_t stream) {
  // the leader variant only exists for check() (no positions)
  auto kern = (args.checkLeader && !WANT_START)
      ? batch_lanes_kernel<STYLE, false, SMEM, K, ClsT, true>
      : batch_lanes_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
  size_t smem = kLanesSmemExtra + (SMEM ? pd.tableBytes : pd.hotBytes);
  cudaError_t err = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess)
    return err;
  int threads = tune.threads ? tune.threads : kMaxThreads;
  int perSm = tune.ctasPerSm ? tune.ctasPerSm : 1;
  uint64_t warps = static_cast<uint64_t>(geo.smCount) * perSm * (threads / 32);
  // Lines per lane per claimed block.  Measured on B200 (profiles/): small
  // blocks win - adjacent lanes then read adjacent lines at about the same
  // time, so the 32-byte sectors they share are still in L1.
  (void)warps;
  uint64_t T = tune.linesPerLane ? tune.linesPerLane : 2;
  T = std::max<uint64_t>(1, std::min<uint64_t>(T, 64));
  uint64_t blocks = (args.nLines + 32 * T - 1) / (32 * T);
  uint64_t ctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * perSm,
                                     (blocks + threads / 32 - 1) / (threads / 32));
  if (ctas == 0)
    return cudaSuccess;
  err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
  if (err != cudaSuccess)
    return err;
  // tiles per group of the sorted schedule: what fits beside the table, and
  // not so many that a small batch leaves SMs without a group
  uint32_t tiles = 0;
  // Measured on B200 (profiles/r1_sorted_sweep.txt): sorting pays where a
  // step is expensive (positions tracked: +14-17%); for the plain styles it
  // gains 1-5% but the scattered tiles cost L1/L2 locality (DRAM traffic
  // x1.7), so those stay on the plain tile schedule.
  if (((tune.variant == 0) && WANT_START) || (tune.variant == 4)) {
    const size_t room = (static_cast<size_t>(geo.maxSmemOptin) > smem)
        ? static_cast<size_t>(geo.maxSmemOptin) - smem : 0;
    uint32_t want = tune.linesPerLane ? static_cast<uint32_t>(tune.linesPerLane) : 8;
    want = std::max(1u, std::min(want, 32u));
    const uint64_t groupsWanted = static_cast<uint64_t>(geo.smCount) * perSm * (threads / 32) * 2;
    while (!tune.linesPerLane && (want > 1) &&
           ((args.nLines + 32ull * want - 1) / (32ull * want) < groupsWanted))
      want >>= 1;
    while ((want > 1) && (sorted_smem_bytes(threads / 32, want) > room))
      want >>= 1;
    if (sorted_smem_bytes(threads / 32, want) <= room)
      tiles = want;
  }
  if (tiles) {
    auto sk = (args.checkLeader && !WANT_START)
        ? batch_sorted_kernel<STYLE, false, SMEM, K, ClsT, true>
        : batch_sorted_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
    // the sort scratch sits behind the tables (+256: worst-case alignment pad)
    const uint32_t sortOffset = static_cast<uint32_t>((kClsBytes + (SMEM ? pd.tableBytes : pd.hotBytes) + 15) & ~15u);
    const size_t total = smem + sorted_smem_bytes(threads / 32, tiles) + 16;
    err = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(total));
    if (err != cudaSuccess)
      return err;
    uint64_t groups = (args.nLines + 32ull * tiles - 1) / (32ull * tiles);
    uint64_t sctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * perSm,
                                        (groups + threads / 32 - 1) / (threads / 32));
    sk<<<static_cast<unsigned>(sctas), threads, total, stream>>>(pd, args, counter, tiles, sortOffset);
    return cudaGetLastError();
  }
  if (tune.variant == 5) {
    auto tk = (args.checkLeader && !WANT_START)
        ? batch_tile32_kernel<STYLE, false, SMEM, K, ClsT, true>
        : batch_tile32_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
    err = cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess)
      return err;
    uint64_t tiles32 = (args.nLines + 31) / 32;
    uint64_t tctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * perSm,
                                        (tiles32 + threads / 32 - 1) / (threads / 32));
    tk<<<static_cast<unsigned>(tctas), threads, smem, stream>>>(pd, args, counter);
    return cudaGetLastError();
  }
  if (tune.variant != 3) {
    auto tk = (args.checkLeader && !WANT_START)
        ? batch_tile_kernel<STYLE, false, SMEM, K, ClsT, true>
        : batch_tile_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
    err = cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess)
      return err;
    uint64_t tiles = (args.nLines + 31) / 32;
    uint64_t tctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * perSm,
                                        (tiles + threads / 32 - 1) / (threads / 32));
    tk<<<static_cast<unsigned>(tctas), threads, smem, stream>>>(pd, args, counter);
    return cudaGetLastError();
  }
  kern<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(
      pd, args, counter, static_cast<uint32_t>(T));
  return cudaGetLastError();
}

template <int STYLE, bool WANT_START>
cudaError_t lanes_table(const ProgramDev &pd, bool smemTable,
                        const BatchArgs &args, const LaunchGeometry &geo,
                        const BatchTuning &tune, unsigned long long *counter,
                        cudaStream_t stream) {
  const bool cls8 = pd.clsMaxOffset < 256;
  if (smemTable) {
    if (pd.unitShift == 0)
      return cls8 ? launch_lanes<STYLE, WANT_START, true, 1, uint8_t>(pd, args, geo, tune, counter, stream)
                  : launch_lanes<STYLE, WANT_START, true, 1, uint16_t>(pd, args, geo, tune, counter, stream);
    if (pd.unitShift == 1)
      return cls8 ? launch_lanes<STYLE, WANT_START, true, 2, uint8_t>(pd, args, geo, tune, counter, stream)
                  : launch_lanes<STYLE, WANT_START, true, 2, uint16_t>(pd, args, geo, tune, counter, stream);
    if (pd.unitShift == 8)  // byte-indexed rows: one load per byte
      return launch_lanes<STYLE, WANT_START, true, 9, uint16_t>(pd, args, geo, tune, counter, stream);
    return cudaErrorInvalidValue;
  }
  return cls8 ? launch_lanes<STYLE, WANT_START, false, 0, uint8_t>(pd, args, geo, tune, counter, stream)
              : launch_lanes<STYLE, WANT_START, false, 0, uint16_t>(pd, args, geo, tune, counter, stream);
}

template <bool WANT_START>
cudaError_t lanes_style(const ProgramDev &pd, bool smemTable,
                        const BatchArgs &args, int style,
                        const LaunchGeometry &geo, const BatchTuning &tune,
                        unsigned long long *counter, cudaStream_t stream) {
  switch (style) {
  case STY_INSTANT: return lanes_table<STY_INSTANT, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_FIRST:   return lanes_table<STY_FIRST, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_TANGENT: return lanes_table<STY_TANGENT, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_LAST:    return lanes_table<STY_LAST, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_FULL:    return lanes_table<STY_FULL, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  default:          return cudaErrorInvalidValue;
  }
}

} // anonymous

cudaError_t launch_batch_lanes(const ProgramDev &pd, bool smemTable,
                               const BatchArgs &args, int style,
                               const LaunchGeometry &geo,
                               const BatchTuning &tune,
                               unsigned long long *counter,
                               cudaStream_t stream, int *launches) {
  cudaError_t err = args.matchStarts
      ? lanes_style<true>(pd, smemTable, args, style, geo, tune, counter, stream)
      : lanes_style<false>(pd, smemTable, args, style, geo, tune, counter, stream);
  if ((err == cudaSuccess) && launches && args.nLines)
    *launches += 1;
  return err;
}

} // namespace red_b200
