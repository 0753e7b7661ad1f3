/* kernels_lanes.cu - the entry point of the check/match batch kernels
   (launch_batch_lanes, at the end of this file: it picks the schedule) and the
   persistent-lane schedule, one of the three that share the chunk engine
   (chunk_engine.cuh):
     kernels_sorted.cu  (default)              tiles of equally long lines
     kernels_tile.cu    (RED_B200_VARIANT=2)   tiles of 32 consecutive lines
     this file          (RED_B200_VARIANT=3)   persistent lanes, described next
   Measured against each other on B200 in profiles/; every schedule is
   bit-exact (tests/test_gpu_parity.py::test_kernel_variants_agree).

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

#include "lanes_dispatch.cuh"

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

struct LanesLauncher {
  template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT>
  static cudaError_t go(const ProgramDev &pd, const BatchArgs &args, const LaunchGeometry &geo,
                        const BatchTuning &tune, unsigned long long *counter, cudaStream_t stream) {
    // the leader variant only exists for check() (no positions)
    auto kern = (args.checkLeader && !WANT_START)
        ? batch_lanes_kernel<STYLE, false, SMEM, K, ClsT, true>
        : batch_lanes_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
    const LaunchShape sh = launch_shape<SMEM>(pd, tune);
    cudaError_t err = cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sh.smem));
    if (err != cudaSuccess)
      return err;
    // Lines per lane per claimed block.  Measured on B200 (profiles/): small
    // blocks win - adjacent lanes then read adjacent lines at about the same
    // time, so the 32-byte sectors they share are still in L1.
    uint64_t T = tune.linesPerLane ? tune.linesPerLane : 2;
    T = std::max<uint64_t>(1, std::min<uint64_t>(T, 64));
    const uint64_t blocks = (args.nLines + 32 * T - 1) / (32 * T);
    const uint64_t ctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * sh.perSm,
                                             (blocks + sh.threads / 32 - 1) / (sh.threads / 32));
    kern<<<static_cast<unsigned>(ctas), sh.threads, sh.smem, stream>>>(
        pd, args, counter, static_cast<uint32_t>(T));
    return cudaGetLastError();
  }
};

} // anonymous

cudaError_t launch_family_lanes(const ProgramDev &pd, bool smemTable, const BatchArgs &args, int style,
                                const LaunchGeometry &geo, const BatchTuning &tune,
                                unsigned long long *counter, cudaStream_t stream) {
  return dispatch_family<LanesLauncher>(pd, smemTable, args, style, geo, tune, counter, stream);
}

/* Entry point for check/match batches: picks the schedule.
     variant 0 (default)  length-sorted tiles (kernels_sorted.cu: a pre-sort
                          kernel, then the batch kernel; variant 4 sorts per
                          warp inside the batch kernel instead); with start
                          positions in two phases (starts only for matched
                          lines).  Plain tiles when the sort scratch does not
                          fit beside the table.
     variant 2            plain tiles (start tracking in every step)
     variant 3            persistent lanes
   Measured: profiles/r1_sorted_sweep.txt, r1_start_sweep.txt.             */
cudaError_t launch_batch_lanes(const ProgramDev &pd, bool smemTable,
                               const BatchArgs &args, int style,
                               const LaunchGeometry &geo,
                               const BatchTuning &tune,
                               unsigned long long *counter,
                               cudaStream_t stream, int *launches) {
  cudaError_t err = cudaSuccess;
  int kernels = 0;
  if (args.nLines == 0)
    return cudaSuccess;
  if ((tune.variant == 0) || (tune.variant == 4))
    err = launch_family_sorted(pd, smemTable, args, style, geo, tune, counter, stream, &kernels);
  if ((err == cudaSuccess) && (kernels == 0)) {
    if (tune.variant == 3)
      err = launch_family_lanes(pd, smemTable, args, style, geo, tune, counter, stream);
    else
      err = launch_family_tile(pd, smemTable, args, style, geo, tune, counter, stream);
    kernels = 1;
  }
  if ((err == cudaSuccess) && launches)
    *launches += kernels;
  return err;
}

} // namespace red_b200
