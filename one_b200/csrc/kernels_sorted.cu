/* kernels_sorted.cu - the production schedule of the check/match batch
   kernels: tiles of 32 lines that span the same number of 16-byte chunks.

   block_sort_kernel    sorts the line numbers of every block of 4096 lines by
                        chunk count and end alignment (2 bytes per line written);
   batch_sorted_kernel  takes tiles of 32 consecutive entries of that array,
                        walks them in step with the chunk engine
                        (chunk_engine.cuh), a tight mask-free loop between the
                        head and the tail chunk, and - when start positions are
                        asked for - finds them in a second phase, for matched
                        lines only.
   RED_B200_VARIANT=4 sorts per warp inside the batch kernel instead (the first
   version, kept for A/B).  DESIGN.md has the measurements. */

#include "lanes_dispatch.cuh"
#include "stream.h"

namespace red_b200 {

namespace {

using namespace dev;

/* batch_sorted_kernel: the tile schedule with tiles of SIMILAR lines.  The
   paragraph below describes the per-warp sort (RED_B200_VARIANT=4); with the
   pre-sorted array (gOrder, the default) a group is one tile of that array.
   In a tile every lane waits for the longest line; with consecutive
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

constexpr uint32_t kQueue = 64;  // start-position queue entries per warp (< 32 waiting + <= 32 new)
inline size_t queue_smem_bytes(int warps) { return static_cast<size_t>(warps) * kQueue * 12; }

// bytes of shared memory the sort needs per CTA of `warps` warps
inline size_t sorted_smem_bytes(int warps, uint32_t T) {
  return static_cast<size_t>(warps) * (32u * T * 2u + 32u * 4u);
}

/* Pre-sort (the default): a small kernel sorts the line numbers of every block
   of kSortBlock consecutive lines by chunk count and writes them, relative to
   the block, as u16 (2 bytes per line of extra traffic).  The batch kernel then
   takes its tiles as 32 consecutive entries of that array: with ~400 lines per
   bucket and block nearly every tile is of one chunk count, where the
   per-warp sort of 128 lines spreads a tile over two or three.  Tiles are
   handed out in order, so the tiles in flight at any time come from a few
   dozen adjacent blocks: neighbouring lines are walked at about the same time
   and the DRAM traffic stays at the algorithmic minimum. */
// (blocks of 16384 lines, where a tile is of one bucket in all three parts of
// the key, measured 2% slower: the sort itself takes longer)
constexpr uint32_t kSortBlock = 4096;
constexpr int kSortThreads = 256;
// chunk count (32) x leading empty quarters of the first chunk (4) x trailing
// empty quarters of the last chunk (4): the lanes of a tile then also agree on
// which quarters of the head and tail chunks hold no byte of anybody's line
constexpr uint32_t kSortBuckets = 512;

__global__ void __launch_bounds__(kSortThreads)
block_sort_kernel(const BatchArgs args, uint16_t *order) {
  __shared__ uint32_t hist[kSortBuckets];
  __shared__ uint16_t sorted[kSortBlock];
  constexpr int kPer = kSortBlock / kSortThreads;  // lines per thread, strided: coalesced loads
  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint64_t nBlocks = (args.nLines + kSortBlock - 1) / kSortBlock;
  const uintptr_t textLow = reinterpret_cast<uintptr_t>(args.text) & 15;
  for (uint64_t blk = blockIdx.x; blk < nBlocks; blk += gridDim.x) {
    const uint64_t base = blk * kSortBlock;
    const uint32_t cnt = (base < n) ? static_cast<uint32_t>(min(n - base, static_cast<uint64_t>(kSortBlock))) : 0;
    for (uint32_t b = threadIdx.x; b < kSortBuckets; b += kSortThreads)
      hist[b] = 0;
    // all index loads of the thread are issued before the first is used
    uint32_t len[kPer];
    uint32_t low[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint32_t i = threadIdx.x + kSortThreads * k;
      len[k] = (i < cnt) ? __ldg(args.lens + base + i) : 0;
      low[k] = (i < cnt) ? static_cast<uint32_t>(__ldg(args.starts + base + i)) : 0;
    }
    uint16_t bucket[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      // chunks the line spans from its 16-byte aligned start: what a lane walks
      const uint32_t skip = static_cast<uint32_t>((textLow + low[k]) & 15);
      const uint32_t chunks = min(31u, (skip + len[k] + 15u) >> 4);
      const uint32_t tailEmpty = (16u - ((skip + len[k]) & 15u)) & 15u;  // bytes past the line in its last chunk
      bucket[k] = static_cast<uint16_t>((chunks * 4 + (skip >> 2)) * 4 + (tailEmpty >> 2));
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (threadIdx.x + kSortThreads * k < cnt)
        atomicAdd(&hist[bucket[k]], 1u);
    __syncthreads();
    if (threadIdx.x < 32) {
      // exclusive scan of the counters: 16 consecutive ones per lane
      constexpr int kEach = kSortBuckets / 32;
      uint32_t sum = 0;
#pragma unroll
      for (int b = 0; b < kEach; ++b)
        sum += hist[threadIdx.x * kEach + b];
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, incl, d);
        if (threadIdx.x >= d)
          incl += up;
      }
      uint32_t run = incl - sum;
#pragma unroll
      for (int b = 0; b < kEach; ++b) {
        const uint32_t c = hist[threadIdx.x * kEach + b];
        hist[threadIdx.x * kEach + b] = run;
        run += c;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint32_t i = threadIdx.x + kSortThreads * k;
      if (i < cnt)
        sorted[atomicAdd(&hist[bucket[k]], 1u)] = static_cast<uint16_t>(i);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint32_t i = threadIdx.x + kSortThreads * k;
      order[base + i] = (i < cnt) ? sorted[i] : static_cast<uint16_t>(kNoIdx);
    }
    __syncthreads();
  }
}

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT, bool LEADER>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_sorted_kernel(const ProgramDev pd, const BatchArgs args,
                    unsigned long long *groupCounter, const uint32_t T,
                    const uint32_t sortOffset, const uint32_t queueOffset,
                    const uint16_t *gOrder) {
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

  // WANT_START: two phases.  The walk itself never tracks the start position
  // (three more instructions in every step of every line); it notes how many
  // bytes the reference's loop would have consumed, and the lines that matched
  // are queued per warp; whenever 32 of them wait (and at the end) each lane
  // takes one and walks its consumed bytes again, keeping only the last
  // position at which the initial state was left (Matcher.h:436-439).
  // Unmatched lines - the bulk of a log - pay nothing.
  // (line, bytes consumed) of matched lines waiting for their start position
  uint64_t *qLine = reinterpret_cast<uint64_t *>(smem + queueOffset) + warp * kQueue;
  uint32_t *qLen = reinterpret_cast<uint32_t *>(smem + queueOffset + nWarps * kQueue * 8) + warp * kQueue;
  uint32_t qCount = 0;

  // phase 2 for entries [base, base + cnt) of the queue, one per lane: walk
  // the consumed part of the line again, keeping only the last position at
  // which the initial state was left (Matcher.h:436-439)
  auto drain = [&](uint32_t base, uint32_t cnt) {
    bool act = lane < cnt;
    const uint64_t dl = act ? qLine[base + lane] : 0;
    const uint32_t used = act ? qLen[base + lane] : 0;
    const uint8_t *dp = args.text + (act ? __ldg(args.starts + dl) : 0);
    const uintptr_t da = reinterpret_cast<uintptr_t>(dp);
    const uint32_t dskip = act ? static_cast<uint32_t>(da & 15) : 0;
    const uint4 *dblk = reinterpret_cast<const uint4 *>(da - dskip);
    const uint32_t dtotal = dskip + used;
    const uint32_t eInit = fast.enc(pd.initV);
    uint32_t e = eInit, ms = 0;
    bool live = act && (used > 0);
    uint4 dcur = make_uint4(0, 0, 0, 0), dnxt = dcur;
    if (live)
      dcur = __ldg(dblk);
    uint32_t dk = 0;
    while (__any_sync(kFull, live)) {
      uint32_t lo = 0, hi = 0;
      bool lastChunk = true;
      if (live) {
        const uint32_t rem = dtotal - dk * 16;
        lo = (dk == 0) ? dskip : 0;
        hi = min(16u, rem);
        lastChunk = rem <= 16;
        if (!lastChunk)
          dnxt = __ldg(dblk + dk + 1);
      }
      const uint32_t mask = (0xffffu >> (16 - hi)) & (0xffffu << lo);
      const uint32_t pos0 = dk * 16 - dskip;
      if (mask != 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool on = (mask & (1u << j)) != 0;
          const uint32_t ne = fast.step(e, word_of(dcur, j >> 2), j & 3);
          if (on && (e == eInit) && (ne != eInit))
            ms = pos0 + j;
          if (on)
            e = ne;
        }
      }
      if (lastChunk)
        live = false;
      ++dk;
      dcur = dnxt;
    }
    if (act)
      args.matchStarts[dl] = ms;
  };


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
  // bucket = chunks the line spans from its 16-byte aligned start (what a lane
  // will actually walk); lines of one bucket finish in the same iteration
  auto bucketOf = [&](uint64_t i) -> uint32_t {
    const uint32_t len = __ldg(args.lens + i);
    const uint32_t skip = static_cast<uint32_t>(
        (reinterpret_cast<uintptr_t>(args.text) + __ldg(args.starts + i)) & 15);
    return min(31u, (skip + len + 15u) >> 4);
  };
  // counting sort of the group's lines by chunk count -> order[]
  auto prepare = [&](uint64_t base) {
    const uint32_t cnt = static_cast<uint32_t>(min(n - base, perGroup));
    hist[lane] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < cnt; i += 32)
      atomicAdd(&hist[bucketOf(base + i)], 1u);
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
      const uint32_t pos = atomicAdd(&hist[bucketOf(base + i)], 1u);
      order[pos] = static_cast<uint16_t>(i);
    }
    for (uint32_t i = cnt + lane; i < perGroup; i += 32)
      order[i] = kNoIdx;
    __syncwarp();
  };

  // in pre-sorted mode a block's entries sit at its front: the last block is
  // walked up to its last (partial) tile
  uint64_t nEnd = n;
  if (gOrder && n) {
    const uint64_t lastBase = (n - 1) / kSortBlock * kSortBlock;
    nEnd = lastBase + ((n - lastBase + 31) & ~uint64_t(31));
  }
  // pre-sorted mode (gOrder): T is 1, a group is one tile of the sorted array;
  // its entries are loaded a tile ahead
  auto loadOrder = [&](uint64_t base) -> uint32_t {
    return (gOrder && (base < nEnd)) ? __ldg(gOrder + base + lane) : kNoIdx;
  };
  uint64_t gBase = claimGroup();
  uint64_t gNext = claimGroup();
  uint32_t idxNow = loadOrder(gBase), idxNext = loadOrder(gNext);
  uint32_t t = 0;
  bool more = gBase < nEnd;
  if (more && !gOrder)
    prepare(gBase);
  if (!gOrder)
    prefetchGroup(gNext);

  // next tile: its line for this lane (~0: none), with the first chunk
  uint64_t nLine = ~0ull;
  const uint8_t *nPtr = nullptr;
  uint32_t nLen = 0;
  uint4 nFirst = make_uint4(0, 0, 0, 0);
  auto pickTile = [&]() {
    if (gOrder) {
      nLine = (idxNow != kNoIdx) ? (gBase & ~static_cast<uint64_t>(kSortBlock - 1)) + idxNow : ~0ull;
      if (nLine >= n)
        nLine = ~0ull;
    }
    else {
      const uint32_t idx = order[t * 32 + lane];
      nLine = (idx != kNoIdx) ? gBase + idx : ~0ull;
    }
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
    Walk<STYLE, false> wk;
    wk.begin(pd);
    bool noResult = false;
    bool ranOut = true;  // the walk consumed the whole line
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
      idxNow = idxNext;
      more = gBase < nEnd;
      if (more) {
        if (!gOrder)
          prepare(gBase);
        gNext = claimGroup();
        if (gOrder)
          idxNext = (gNext < nEnd) ? __ldg(gOrder + gNext + lane) : kNoIdx;
        else
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
    // one iteration with everything a partial chunk needs (head, tail, lines
    // of different length in one tile)
    auto genericStep = [&]() {
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
      const bool allFull = __all_sync(kFull, (mask == 0xffffu) || (mask == 0));
      // quarters of the chunk in which some lane has a byte
      uint32_t quarters = 0xfu;
      if (!allFull)
        quarters = (__any_sync(kFull, mask & 0x000fu) ? 1u : 0u) | (__any_sync(kFull, mask & 0x00f0u) ? 2u : 0u) |
                   (__any_sync(kFull, mask & 0x0f00u) ? 4u : 0u) | (__any_sync(kFull, mask & 0xf000u) ? 8u : 0u);
      const bool stopped = process_chunk<STYLE, false>(pd, fast, tab, cls, cur, mask, pos0, wk, active,
                                                       allFull, quarters);
      if (stopped)
        ranOut = false;
      if (stopped || lastChunk)
        active = false;
      ++k;
      cur = nxt;
    };
    // The tile's lines span (nearly) the same number of chunks.  Chunks
    // 1 .. kMin-2 are whole for every lane still walking, so between the head
    // and the tail runs a loop without masks, votes or per-lane bookkeeping.
    const uint32_t kMin = __reduce_min_sync(kFull, active ? ((total + 15u) >> 4) : 0xffffffffu);
    if (__any_sync(kFull, active))
      genericStep();  // the head chunk
    if ((kMin != 0xffffffffu) && (kMin >= 3)) {
      if (wantFirst) {
        fetchFirst();
        wantFirst = false;
      }
      for (; k + 1 < kMin; ++k) {
        if (active)
          nxt = __ldg(blk + k + 1);
        const bool stopped = process_chunk<STYLE, false>(
            pd, fast, tab, cls, cur, active ? 0xffffu : 0u, k * 16 - skip, wk, active, true);
        if (stopped) {
          ranOut = false;
          active = false;
        }
        cur = nxt;
      }
    }
    while (__any_sync(kFull, active))
      genericStep();
    if (wantFirst)
      fetchFirst();

    int32_t result = 0;
    if (line != ~0ull) {
      uint32_t start = 0, end = 0;
      const uint32_t consumed = ranOut ? lineLen : wk.mstart;
      if (!(LEADER && noResult))
        wk.finish(pd, lineLen, result, start, end);
      args.results[line] = result;
      if (WANT_START && (result <= 0))
        args.matchStarts[line] = 0;
      if (args.matchEnds)
        args.matchEnds[line] = end;
      wk.mstart = consumed;  // queued below
    }
    if (WANT_START) {
      const bool hit = (line != ~0ull) && (result > 0);
      const unsigned hits = __ballot_sync(kFull, hit);
      if (hit) {
        const uint32_t at = qCount + __popc(hits & ((1u << lane) - 1));
        qLine[at] = line;
        qLen[at] = wk.mstart;
      }
      qCount += __popc(hits);
      __syncwarp();
      if (qCount >= 32) {
        drain(qCount - 32, 32);
        qCount -= 32;
        __syncwarp();
      }
    }
  }
  if (WANT_START && qCount)
    drain(0, qCount);
}

// set by SortedLauncher::go (per host thread): kernels it launched (0: it did not fit)
thread_local int gSortedKernels = 0;

struct SortedLauncher {
  template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT>
  static cudaError_t go(const ProgramDev &pd, const BatchArgs &args, const LaunchGeometry &geo,
                        const BatchTuning &tune, unsigned long long *counter, cudaStream_t stream) {
    const LaunchShape sh = launch_shape<SMEM>(pd, tune);
    const int warps = sh.threads / 32;
    const size_t queueBytes = WANT_START ? queue_smem_bytes(warps) + 16 : 0;
    const size_t room = (static_cast<size_t>(geo.maxSmemOptin) > sh.smem + 16 + queueBytes)
        ? static_cast<size_t>(geo.maxSmemOptin) - sh.smem - 16 - queueBytes : 0;
    // variant 4: the per-warp sort inside the kernel, T tiles per group.
    // Measured on B200, 20 M lines (profiles/r1_sorted_T_traffic.txt):
    //   T  GB/s   DRAM bytes / algorithmic
    //   2  1298   1.01
    //   4  1369   1.13
    //   8  1396   1.71   (148 x 32 groups of 26 KB outlive the L2: the 64-byte
    //                     pieces two neighbouring lines share are fetched twice)
    const bool presort = (tune.variant != 4);
    uint32_t tiles = 1;
    if (!presort) {
      uint32_t want = tune.linesPerLane ? static_cast<uint32_t>(tune.linesPerLane) : 4;
      want = std::max(1u, std::min(want, 32u));
      const uint64_t groupsWanted = static_cast<uint64_t>(geo.smCount) * sh.perSm * warps * 2;
      while (!tune.linesPerLane && (want > 1) &&
             ((args.nLines + 32ull * want - 1) / (32ull * want) < groupsWanted))
        want >>= 1;
      while ((want > 1) && (sorted_smem_bytes(warps, want) > room))
        want >>= 1;
      tiles = want;
    }
    const size_t sortBytes = presort ? 0 : sorted_smem_bytes(warps, tiles);
    if (sortBytes > room)
      return cudaSuccess;  // does not fit: the caller falls back to plain tiles
    if (WANT_START && (room == 0))
      return cudaSuccess;
    auto sk = (args.checkLeader && !WANT_START)
        ? batch_sorted_kernel<STYLE, false, SMEM, K, ClsT, true>
        : batch_sorted_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
    // the sort scratch sits behind the tables, the start queue behind that
    const uint32_t sortOffset = static_cast<uint32_t>((kClsBytes + (SMEM ? pd.tableBytes : pd.hotBytes) + 15) & ~15u);
    const uint32_t queueOffset = static_cast<uint32_t>((sortOffset + sortBytes + 15) & ~size_t(15));
    const size_t total = sh.smem + sortBytes + 16 + queueBytes;
    cudaError_t err = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(total));
    if (err != cudaSuccess)
      return err;
    uint16_t *order = nullptr;
    if (presort) {
      const uint64_t nBlocks = (args.nLines + kSortBlock - 1) / kSortBlock;
      err = scratch_alloc(&order, nBlocks * kSortBlock * sizeof(uint16_t), stream);
      if (err != cudaSuccess)
        return err;
      const unsigned sgrid = static_cast<unsigned>(std::min<uint64_t>(nBlocks, static_cast<uint64_t>(geo.smCount) * 8));
      block_sort_kernel<<<sgrid, kSortThreads, 0, stream>>>(args, order);
      gSortedKernels += 1;
    }
    const uint64_t groups = (args.nLines + 32ull * tiles - 1) / (32ull * tiles);
    const uint64_t ctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * sh.perSm,
                                             (groups + warps - 1) / warps);
    sk<<<static_cast<unsigned>(ctas), sh.threads, total, stream>>>(pd, args, counter, tiles, sortOffset,
                                                                  queueOffset, order);
    gSortedKernels += 1;
    err = cudaGetLastError();
    scratch_free(order, stream);
    return err;
  }
};

} // anonymous

cudaError_t launch_family_sorted(const ProgramDev &pd, bool smemTable, const BatchArgs &args, int style,
                                 const LaunchGeometry &geo, const BatchTuning &tune,
                                 unsigned long long *counter, cudaStream_t stream, int *kernels) {
  *kernels = 0;
  if (args.nLines == 0)
    return cudaSuccess;
  gSortedKernels = 0;
  cudaError_t err = dispatch_family<SortedLauncher>(pd, smemTable, args, style, geo, tune, counter, stream);
  *kernels = gSortedKernels;
  return err;
}

} // namespace red_b200
