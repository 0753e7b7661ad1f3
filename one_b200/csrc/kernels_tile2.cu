/* kernels_tile2.cu - match() WITH start positions, in two phases.

   Tracking "where was the initial state last left" (Outcome::start_,
   Matcher.h:436-439) costs three more instructions in every step of every
   line, and the step is what bounds the kernel (ncu: profiles/): the
   single-phase kernels run at 0.65-0.74 TB/s with it against 1.2 TB/s without.
   But the start position only means something for a line that matched.  So:

     phase 1  the plain tile schedule without start tracking; besides the
              outcome it notes how many bytes the reference's loop would have
              consumed (where the walk stopped);
     phase 2  matched lines are queued per warp in shared memory; whenever 32
              of them wait (and at the end), each lane takes one and walks its
              consumed bytes again with a loop that does nothing but keep the
              last position at which the initial state was left.

   Unmatched lines - the bulk of a log - pay nothing. */

#include "lanes_dispatch.cuh"

namespace red_b200 {

namespace {

using namespace dev;

constexpr uint32_t kQueue = 64;  // queue entries per warp (< 32 waiting + <= 32 new)

inline size_t queue_smem_bytes(int warps) { return static_cast<size_t>(warps) * kQueue * 12; }

template <int STYLE, bool SMEM, int K, class ClsT>
__global__ void __launch_bounds__(kMaxThreads, 1)
batch_tile2_kernel(const ProgramDev pd, const BatchArgs args,
                   unsigned long long *blockCounter, const uint32_t queueOffset) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  using Eng = Engine<SMEM, K, ClsT>;
  unsigned char *smem = Eng::stage(pd, smemRaw);
  const ClsT *cls = reinterpret_cast<const ClsT *>(smem);
  const typename Eng::Tab tab = Eng::table(pd, smem);
  const typename Eng::Fast fast = Eng::fast(pd, smem);

  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint32_t leadLen = 0;
  const uint32_t warp = threadIdx.x >> 5;
  // matched lines waiting for their start position: (line, bytes consumed)
  uint64_t *qLine = reinterpret_cast<uint64_t *>(smem + queueOffset) + warp * kQueue;
  uint32_t *qLen = reinterpret_cast<uint32_t *>(smem + queueOffset + (blockDim.x >> 5) * kQueue * 8) + warp * kQueue;
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
    Walk<STYLE, false> wk;
    wk.begin(pd);
    bool ranOut = true;  // the walk consumed the whole line
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
          process_chunk<STYLE, false>(pd, fast, tab, cls, cur, mask, pos0, wk, active);
      if (stopped)
        ranOut = false;
      if (stopped || lastChunk)
        active = false;
      ++k;
      cur = nxt;
    }
    if (wantFirst)
      fetchFirst();

    int32_t result = 0;
    if (line < n) {
      uint32_t start = 0, end = 0;
      const uint32_t consumed = ranOut ? lineLen : wk.mstart;
      wk.finish(pd, lineLen, result, start, end);
      (void)start;
      args.results[line] = result;
      if (result <= 0)
        args.matchStarts[line] = 0;
      if (args.matchEnds)
        args.matchEnds[line] = end;
      wk.mstart = consumed;  // queued below
    }
    // queue the matched lines; once a full warp of them waits, find their starts
    const bool hit = (line < n) && (result > 0);
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
  if (qCount)
    drain(0, qCount);
}

struct Tile2Launcher {
  template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT>
  static cudaError_t go(const ProgramDev &pd, const BatchArgs &args, const LaunchGeometry &geo,
                        const BatchTuning &tune, unsigned long long *counter, cudaStream_t stream) {
    if constexpr (!WANT_START) {
      (void)pd; (void)args; (void)geo; (void)tune; (void)counter; (void)stream;
      return cudaErrorInvalidValue;  // only used with start positions
    }
    else {
      auto kern = batch_tile2_kernel<STYLE, SMEM, K, ClsT>;
      const LaunchShape sh = launch_shape<SMEM>(pd, tune);
      const uint32_t queueOffset = static_cast<uint32_t>((kClsBytes + (SMEM ? pd.tableBytes : pd.hotBytes) + 15) & ~15u);
      const size_t total = sh.smem + queue_smem_bytes(sh.threads / 32) + 16;
      if (total > static_cast<size_t>(geo.maxSmemOptin))
        return cudaErrorInvalidConfiguration;  // the caller falls back
      cudaError_t err = cudaFuncSetAttribute(
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(total));
      if (err != cudaSuccess)
        return err;
      const uint64_t tiles = (args.nLines + 31) / 32;
      const uint64_t ctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * sh.perSm,
                                               (tiles + sh.threads / 32 - 1) / (sh.threads / 32));
      kern<<<static_cast<unsigned>(ctas), sh.threads, total, stream>>>(pd, args, counter, queueOffset);
      return cudaGetLastError();
    }
  }
};

} // anonymous

cudaError_t launch_family_tile2(const ProgramDev &pd, bool smemTable, const BatchArgs &args, int style,
                                const LaunchGeometry &geo, const BatchTuning &tune,
                                unsigned long long *counter, cudaStream_t stream) {
  return dispatch_family<Tile2Launcher>(pd, smemTable, args, style, geo, tune, counter, stream);
}

} // namespace red_b200
