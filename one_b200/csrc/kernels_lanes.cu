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

#include "device_program.h"
#include "walk.cuh"

#include <algorithm>
#include <type_traits>

namespace red_b200 {

namespace {

using namespace dev;

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t r;
  asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t r;
  asm("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}

/* Fast-pass policies.  The state is carried in an "encoded" form so that the
   address of the next table entry is one add away.

   SmemFast8 (table in shared memory, class offsets fit a byte): the encoded
   state is the shared-memory address of its row; the class table sits on a
   256-byte boundary so that PRMT builds its address from the text byte and
   the table base in one instruction.  One step is
       PRMT, LDS.U8, IADD, LDS.U16, LEA
   GenericFast: plain accessors (global table, or > 128 classes). */
template <int K> struct SmemFast8 {
  uint32_t clsBase;  // shared address, low byte zero
  uint32_t tBase;    // shared address of the transition table
  __device__ __forceinline__ uint32_t enc(uint32_t v) const { return (v << K) + tBase; }
  __device__ __forceinline__ uint32_t dec(uint32_t e) const { return (e - tBase) >> K; }
  __device__ __forceinline__ uint32_t encLimit(uint32_t lim) const { return lim ? enc(lim) : 0; }
  __device__ __forceinline__ uint32_t step(uint32_t e, uint32_t word, int jj) const {
    const uint32_t a = __byte_perm(word, clsBase, 0x7650 | jj);
    return enc(lds_u16(e + lds_u8(a)));
  }
};

/* SmemFastByte (small DFAs): rows of 256 u16 entries indexed by the raw byte;
   the encoded state is the row's shared address.  One step is
       PRMT, LEA, LDS.U16, LEA  - a single table load per byte. */
struct SmemFastByte {
  uint32_t tBase;
  __device__ __forceinline__ uint32_t enc(uint32_t v) const { return (v << 9) + tBase; }
  __device__ __forceinline__ uint32_t dec(uint32_t e) const { return (e - tBase) >> 9; }
  __device__ __forceinline__ uint32_t encLimit(uint32_t lim) const { return lim ? enc(lim) : 0; }
  __device__ __forceinline__ uint32_t step(uint32_t e, uint32_t word, int jj) const {
    return enc(lds_u16(e + (byte_of(word, jj) << 1)));
  }
};

template <class Tab, class ClsT> struct GenericFast {
  Tab tab;
  const ClsT *cls;
  __device__ __forceinline__ uint32_t enc(uint32_t v) const { return v; }
  __device__ __forceinline__ uint32_t dec(uint32_t e) const { return e; }
  __device__ __forceinline__ uint32_t encLimit(uint32_t lim) const { return lim; }
  __device__ __forceinline__ uint32_t step(uint32_t e, uint32_t word, int jj) const {
    return tab.next(e, cls[byte_of(word, jj)]);
  }
};

/* One aligned 16-byte chunk of one line per lane.  `mask` marks the bytes that
   belong to the line.  Returns true when the walk is over (style decision or a
   state that can never be left); wk carries the walk across chunks. */
template <int STYLE, bool WANT_START, class Fast, class Tab, class ClsT>
__device__ __forceinline__ bool
process_chunk(const ProgramDev &pd, const Fast &fast, const Tab &tab, const ClsT *cls,
              const uint4 &cur, const uint32_t mask, const uint32_t pos0,
              Walk<STYLE, WANT_START> &wk, const bool cValid) {
  bool stopped = false;
  // A lane with nothing in this chunk (its line is over, or it has no line)
  // stays out altogether: its lookups would be pure noise in the shared-memory
  // banks (ncu: 2.6 wavefronts per table load with idle lanes in, see profiles/).
  if (mask != 0) {
        // ---- fast pass: no decisions, quarter maxima only -------------------
        // Within an active lane the two table loads are issued for all 16
        // bytes, also those outside the line (they read valid table memory and
        // the value is dropped): a select is cheaper than 16 branches.
        const uint32_t eIn = fast.enc(wk.v);
        const uint32_t eInit = fast.enc(pd.initV);
        uint32_t e = eIn;
        uint32_t mq0 = 0, mq1 = 0, mq2 = 0, mq3 = 0;
        uint32_t cp0 = 0, cp1 = 0, cp2 = 0;
        uint32_t msFast = wk.mstart;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t word = word_of(cur, q);
          uint32_t m = 0;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * q + jj;
            const bool act = (mask & (1u << j)) != 0;
            const uint32_t ne = fast.step(e, word, jj);
            if (WANT_START) {
              if (act && (e == eInit) && (ne != eInit))
                msFast = pos0 + j;
            }
            if (act)
              e = ne;
            m = max(m, e);
          }
          if (q == 0) { mq0 = m; cp0 = e; }
          if (q == 1) { mq1 = m; cp1 = e; }
          if (q == 2) { mq2 = m; cp2 = e; }
          if (q == 3) { mq3 = m; }
        }
        const uint32_t eLimit = fast.encLimit(wk.limit);
        const uint32_t ePure = fast.enc(pd.pureV);
        const uint32_t eAbs = fast.enc(pd.absV);
        const uint32_t mAll = max(max(mq0, mq1), max(mq2, mq3));
        if ((mAll < eLimit) || !cValid) {
          wk.v = fast.dec(e);
          if (WANT_START)
            wk.mstart = msFast;
        }
        else {
          // ---- careful pass: from the first flagged quarter on, quarter by
          // quarter, with the reference's per-byte decisions.  Quarters whose
          // maximum stayed below the (current) limit are skipped.
          int q = (mq0 >= eLimit) ? 0 : (mq1 >= eLimit) ? 1 : (mq2 >= eLimit) ? 2 : 3;
          if (WANT_START)
            q = 0;
          constexpr bool kLean = !WANT_START &&
              ((STYLE == STY_INSTANT) || (STYLE == STY_LAST) || (STYLE == STY_FULL));
          uint32_t ee = e;
#pragma unroll 1
          for (; (q < 4) && !stopped; ++q) {
            const uint32_t m = (q == 0) ? mq0 : (q == 1) ? mq1 : (q == 2) ? mq2 : mq3;
            if (WANT_START || (m >= fast.encLimit(wk.limit))) {
              const uint32_t eq = (q == 0) ? eIn : (q == 1) ? cp0 : (q == 2) ? cp1 : cp2;
              const uint32_t word = (q == 0) ? cur.x : (q == 1) ? cur.y : (q == 2) ? cur.z : cur.w;
              const uint32_t nib = mask >> (4 * q);
              if (kLean) {
                // Instant / Last / Full without start tracking: the decisions
                // are compares on the encoded state (Matcher.h:441-478)
                ee = eq;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  if (!stopped && (nib & (1u << jj))) {
                    ee = fast.step(ee, word, jj);
                    if (ee >= eLimit) {
                      if (ee >= ePure)
                        stopped = true;           // pure dead end
                      else if (STYLE == STY_FULL)
                        stopped = true;           // absorbing accept: outcome known
                      else {
                        wk.accV = fast.dec(ee);
                        wk.mend = pos0 + 4 * q + jj + 1;
                        if (STYLE == STY_INSTANT)
                          stopped = true;
                        else if (ee >= eAbs) {
                          wk.absorbed = true;
                          stopped = true;
                        }
                      }
                    }
                  }
                }
              }
              else {
                wk.v = fast.dec(eq);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  if (!stopped && (nib & (1u << jj)))
                    stopped = wk.step(pd, tab, cls[byte_of(word, jj)], pos0 + 4 * q + jj);
                }
              }
            }
          }
          if (kLean)
            wk.v = fast.dec(stopped ? ee : e);
          else if (!stopped)
            wk.v = fast.dec(e);
        }
      }
  return stopped;
}

/* Shared memory (dynamic): up to 255 bytes of padding so that the class table
   starts on a 256-byte boundary, 512 bytes of class table, then the
   transition table (smem mode). */
constexpr uint32_t kLanesSmemExtra = 256 + kClsBytes;

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT, bool LEADER>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_lanes_kernel(const ProgramDev pd, const BatchArgs args,
                   unsigned long long *blockCounter, const uint32_t T) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  unsigned char *smem = smemRaw + ((256u - (smem_u32(smemRaw) & 255u)) & 255u);
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes
                     : static_cast<const unsigned char *>(pd.table)};
  constexpr bool kPtx = SMEM && (sizeof(ClsT) == 1);
  constexpr bool kByte = SMEM && (K == 9);
  using Fast = typename std::conditional<kByte, SmemFastByte,
      typename std::conditional<kPtx, SmemFast8<K>, GenericFast<Tab, ClsT>>::type>::type;
  Fast fast;
  if constexpr (kByte) {
    fast.tBase = smem_u32(smem + kClsBytes);
  }
  else if constexpr (kPtx) {
    fast.clsBase = smem_u32(smem);
    fast.tBase = smem_u32(smem + kClsBytes);
  }
  else {
    fast.tab = tab;
    fast.cls = cls;
  }

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
  unsigned char *smem = smemRaw + ((256u - (smem_u32(smemRaw) & 255u)) & 255u);
  stage_classes<ClsT>(pd, smem);
  if (SMEM)
    stage_table(pd, smem + kClsBytes);
  __syncthreads();
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  const Tab tab{SMEM ? smem + kClsBytes
                     : static_cast<const unsigned char *>(pd.table)};
  constexpr bool kPtx = SMEM && (sizeof(ClsT) == 1);
  constexpr bool kByte = SMEM && (K == 9);
  using Fast = typename std::conditional<kByte, SmemFastByte,
      typename std::conditional<kPtx, SmemFast8<K>, GenericFast<Tab, ClsT>>::type>::type;
  Fast fast;
  if constexpr (kByte) {
    fast.tBase = smem_u32(smem + kClsBytes);
  }
  else if constexpr (kPtx) {
    fast.clsBase = smem_u32(smem);
    fast.tBase = smem_u32(smem + kClsBytes);
  }
  else {
    fast.tab = tab;
    fast.cls = cls;
  }

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

template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT>
cudaError_t launch_lanes(const ProgramDev &pd, const BatchArgs &args,
                         const LaunchGeometry &geo, const BatchTuning &tune,
                         unsigned long long *counter, cudaStream_t stream) {
  // the leader variant only exists for check() (no positions)
  auto kern = (args.checkLeader && !WANT_START)
      ? batch_lanes_kernel<STYLE, false, SMEM, K, ClsT, true>
      : batch_lanes_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
  size_t smem = kLanesSmemExtra + (SMEM ? pd.tableBytes : 0);
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
