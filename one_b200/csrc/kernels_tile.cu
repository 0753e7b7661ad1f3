/* kernels_tile.cu - the tile schedule of the check/match batch kernel (the
   production default).  The chunk engine and the description of the design
   are in chunk_engine.cuh and kernels_lanes.cu. */

#include "lanes_dispatch.cuh"

namespace red_b200 {

namespace {

using namespace dev;

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

struct TileLauncher {
  template <int STYLE, bool WANT_START, bool SMEM, int K, class ClsT>
  static cudaError_t go(const ProgramDev &pd, const BatchArgs &args, const LaunchGeometry &geo,
                        const BatchTuning &tune, unsigned long long *counter, cudaStream_t stream) {
    // the leader variant only exists for check() (no positions)
    auto kern = (args.checkLeader && !WANT_START)
        ? batch_tile_kernel<STYLE, false, SMEM, K, ClsT, true>
        : batch_tile_kernel<STYLE, WANT_START, SMEM, K, ClsT, false>;
    const LaunchShape sh = launch_shape<SMEM>(pd, tune);
    cudaError_t err = cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sh.smem));
    if (err != cudaSuccess)
      return err;
    const uint64_t tiles = (args.nLines + 31) / 32;
    const uint64_t ctas = std::min<uint64_t>(static_cast<uint64_t>(geo.smCount) * sh.perSm,
                                             (tiles + sh.threads / 32 - 1) / (sh.threads / 32));
    kern<<<static_cast<unsigned>(ctas), sh.threads, sh.smem, stream>>>(pd, args, counter);
    return cudaGetLastError();
  }
};

} // anonymous

cudaError_t launch_family_tile(const ProgramDev &pd, bool smemTable, const BatchArgs &args, int style,
                               const LaunchGeometry &geo, const BatchTuning &tune,
                               unsigned long long *counter, cudaStream_t stream) {
  return dispatch_family<TileLauncher>(pd, smemTable, args, style, geo, tune, counter, stream);
}

} // namespace red_b200
