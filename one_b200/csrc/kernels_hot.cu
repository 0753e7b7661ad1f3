/* kernels_hot.cu - the staged check/match batch kernel (the production path for
   styles Instant, Last and Full) and the visit-count profile kernel.

   What bounded the first production kernel (kernels_sorted.cu; ncu in
   profiles/r1_final_ncu_summary.json) was the LSU data pipe, one 128-byte
   wavefront per cycle per SM: every lane's 16-byte text load from global
   memory was a wavefront of its own (each lane sits in a different line), and
   every byte cost two dependent shared-memory lookups (class, then
   transition).  Here

     * the TEXT goes through shared memory: a producer warp brings the
       contiguous span of a batch of 256 consecutive lines into a ring of
       slots with one bulk-async copy (cp.async.bulk ... mbarrier::
       complete_tx), asked into L2 while the batch is sorted; the consumer
       lanes read their line with LDS.128;
     * a step is ONE lookup in the hot image (program.h: HotImage): up to 255
       states with one-byte codes and byte-indexed rows; lanes that leave the
       hot set continue on the full table in global memory (L2) until they are
       back;
     * the lanes of a tile walk IN STEP BY POSITION IN THEIR LINES: a lane
       re-aligns its text in registers (hot_walk.cuh: Feed), so byte 0 of all
       32 lines is looked up in the same instruction.  Lines of a log look
       alike at equal positions, and lanes in like states on like bytes share
       a shared-memory word instead of colliding in the banks: 2.0 -> 1.4
       wavefronts per lookup on the bench corpus;
     * the length sort (tiles of 32 lines of equal length in bytes, longest
       first; equal length very often means equal shape) happens inside the
       CTA, per batch, by the producer warp, whose index entries are loaded one
       batch ahead; every lane writes its line's outcome straight to global
       memory;
     * styles that end at the first state needing a decision (Instant, Full)
       walk on an image in which such a state STICKS: the fast pass keeps no
       running maximum, the code behind a 16-byte chunk says whether anything
       happened, a flagged lane steps aside, and when the tile is through all
       flagged lanes settle their outcome together - the state itself names
       it, and for Instant a second run over the QUARTER of the chunk in
       which the run met it counts out where it was entered (the codes behind
       the quarters are kept, the quarter's four bytes too);
     * what lies around the lookups is kept off a tile's critical path: the
       next chunk's text is loaded a chunk ahead, the next tile's ticket while
       the current tile is finished off, ticket arithmetic is by
       multiplication, the outcome's fields stay in registers, the dealing's
       ranks come from two warp sums and a match, the producers' span
       arithmetic is in 32 bits ("What a tile costs besides its
       chain").

   Roles in a CTA: warps 0 .. NC-1 are consumers (a tile of 32 lines at a
   time, claimed in order from a shared-memory ticket), the warps behind them
   producers.  Synchronisation is by mbarriers only: full[slot] (producer's
   arrival + the copy's bytes; only the warp with the batch's first tile waits
   there - every piece of the copy wakes the waiters up), ready[slot] (that
   warp's arrival: where the others sleep), empty[slot] (one arrival per
   tile).

   Lines that are not laid out back to back (shuffled or overlapping views,
   a batch whose span exceeds a slot) are gathered: one bulk copy per line,
   as many lines per batch as fit a slot.  A line longer than a slot is walked
   from global memory on the full table. */

#include "hot_walk.cuh"

#include <algorithm>
#include <type_traits>

namespace red_b200 {

namespace {

#ifndef HOT_UNROLL
#define HOT_UNROLL 2
#endif
constexpr int kHotUnroll = HOT_UNROLL;    // whole chunks per trip of the consumers' interior loop
constexpr uint32_t kMaxBatch = 256;       // lines per batch (and per claimed sub-block)
constexpr uint32_t kSortBuckets = 256;    // by the line's length in bytes (longer lines share the first bucket)
constexpr uint32_t kUnstaged = 0xffffu;   // sOff: the line's text is not in the slot
constexpr uint32_t kMaxSlots = 12;
constexpr uint32_t kMaxProducers = 4;

struct HotParams {
  // byte offsets into dynamic shared memory (behind its 256-byte alignment)
  uint32_t oC2v, oBars, oCtl, oProd, oDeal, oMeta, oSlots;
  uint32_t nSlots, slotBytes, batchLines, tilesPerBatch, metaBytes, nProducers, prodBytes;
  uint32_t tpbMagic, slotsMagic;  // floor(2^32 / d) + 1: x / d = umulhi(x, magic) for x < 2^32 / d (tickets and batch numbers of one CTA are far below)
  unsigned long long *dbg;  // null, or 16 cycle counters (HotTuning::debug)
};

// per-slot metadata (offsets within the slot's block)
constexpr uint32_t kMetaIdx = 0;                          // u8[256] line (relative to the batch) at sorted position
constexpr uint32_t kMetaOL = kMetaIdx + kMaxBatch;        // u32[256] text offset in the slot | length << 16
constexpr uint32_t kMetaArrays = kMetaOL + 4 * kMaxBatch; // (one byte and one word per line: two stores and two loads, not three)
constexpr uint32_t kMetaFirst = kMetaArrays;              // u64 first line of the batch
constexpr uint32_t kMetaCnt = kMetaFirst + 8;             // u32 lines in the batch
constexpr uint32_t kMetaSeq = kMetaCnt + 4;               // u32 number of the batch the slot holds
constexpr uint32_t kMetaBytes = kMetaSeq + 4;

// per-producer scratch
constexpr uint32_t kProdMeta = 0;                             // sorted idx/off/len of the batch being prepared
constexpr uint32_t kProdOff = kProdMeta + kMetaArrays;        // u16[256] text offsets by line (gather path)
constexpr uint32_t kProdHist = kProdOff + 2 * kMaxBatch;      // u32[kSortBuckets]
constexpr uint32_t kProdBytes = kProdHist + 4 * kSortBuckets;

/* DEFER (styles whose walk ends at the first state that needs a decision:
   Instant, Full): a lane whose chunk was flagged steps aside, and the careful
   passes of a tile's flagged lanes run together once the tile is through,
   instead of one after the other with the rest of the warp waiting. */
template <int STYLE, bool DEFER, int STRIDE, bool DBG>
__global__ void __launch_bounds__(1024, 1)
batch_hot_kernel(const ProgramDev pd, const HotDev hd, const BatchArgs args, const HotParams hp,
                 unsigned long long *blockCounter) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  // everything sits behind a 256-byte boundary (hot_next wants the table there)
  unsigned char *smem = smemRaw + ((256u - (smem_addr(smemRaw) & 255u)) & 255u);
  const uint32_t sBase = smem_addr(smem);
  // cycle counters (RED_B200_HOT_DEBUG) are an instantiation of their own: as a
  // run-time switch they cost the consumers ten registers of the 64 a thread has
  const bool dbgOn = DBG && (hp.dbg != nullptr);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t P = hp.nProducers;
  const uint32_t nConsumers = (blockDim.x >> 5) - P;
  const uint32_t R = hp.nSlots;
  const uint32_t TPB = hp.tilesPerBatch;
  const uint32_t barsAddr = sBase + hp.oBars;      // full[0..R), empty[0..R), ready[0..R)
  const uint32_t ctlAddr = sBase + hp.oCtl;        // +0 ticket, +4 next batch number, +8 producers done, +12 endQ, +16 batches that have their slot
  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;

  // stage the hot image, set up the barriers
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(hd.table);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (uint32_t i = threadIdx.x; i < (hd.tableBytes >> 4); i += blockDim.x)
      dst[i] = __ldg(src + i);
    uint32_t *c2v = reinterpret_cast<uint32_t *>(smem + hp.oC2v);
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
      const uint32_t v = __ldg(hd.codeToV + i);
      c2v[i] = v;
      // result id of the state behind a code (0 unless accepting)
      c2v[256 + i] = ((i < hd.nHot) && (v >= pd.acceptV) && (v < pd.pureV))
          ? static_cast<uint32_t>(__ldg(pd.results + state_index(pd, v))) : 0u;
    }
    if (threadIdx.x == 0) {
      for (uint32_t s = 0; s < R; ++s) {
        mbar_init(barsAddr + 8 * s, 1);
        mbar_init(barsAddr + 8 * (R + s), TPB);
        mbar_init(barsAddr + 8 * (2 * R + s), 1);
        sts32(sBase + hp.oMeta + s * hp.metaBytes + kMetaSeq, 0xffffffffu);
      }
      sts32(ctlAddr, 0);
      sts32(ctlAddr + 4, 0);
      sts32(ctlAddr + 8, 0);
      sts32(ctlAddr + 12, 0xffffffffu);
      sts32(ctlAddr + 16, 0);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();

  if (warp >= nConsumers) {
    // ======================= producers =======================================
    /* A producer's per-batch work is a chain of dependent shared-memory
       operations behind the consumers' lookups in the same pipe; several
       producers overlap their chains.  A batch gets its number q (it lives in
       slot q % R, consumers take the batches in order of q) when its producer
       has everything ready and only needs a slot to fill. */
    const uint32_t SUB = hp.batchLines;
    const uint64_t textA = reinterpret_cast<uint64_t>(args.text);
    const uint32_t prodA = sBase + hp.oProd + (warp - nConsumers) * hp.prodBytes;
    const uint32_t scrA = prodA + kProdMeta;
    const uint32_t offA = prodA + kProdOff;
    const uint32_t histA = prodA + kProdHist;
    uint32_t *sHist = reinterpret_cast<uint32_t *>(smem + hp.oProd + (warp - nConsumers) * hp.prodBytes + kProdHist);
    uint32_t q = 0;  // the batch being produced
    constexpr uint32_t kPer = kMaxBatch / 32;

    /* Slots are handed out in the order of q: a parity wait can only tell two
       consecutive phases of a barrier apart, so the producer of batch q may
       look at empty[q % R] only after the producer of batch q - R has seen
       ITS phase complete. */
    unsigned long long tAcq = 0, tIdx = 0, nBatches = 0;
    auto acquire = [&]() -> uint32_t {
      const long long t0 = dbgOn ? clock64() : 0;
      if (lane == 0) {
        q = atomicAdd(reinterpret_cast<uint32_t *>(smem + hp.oCtl + 4), 1u);
        while (lds32_volatile(ctlAddr + 16) != q)
          __nanosleep(32);
      }
      q = __shfl_sync(kFullMask, q, 0);
      const uint32_t s = q % R;
      if (q >= R)
        mbar_wait(barsAddr + 8 * (R + s), ((q / R) - 1) & 1);
      __syncwarp();
      if (lane == 0)
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ctlAddr + 16), "r"(q + 1) : "memory");
      if (dbgOn) {
        tAcq += clock64() - t0;
        ++nBatches;
      }
      return s;
    };
    /* The sort key of a line: its length in bytes, longest first (the lanes of
       a tile walk in step from byte 0 of their lines, so a tile takes as long
       as its longest line - and the tile that takes longest is the first to
       be claimed, so the slot is not held up by a late starter).  Lines of
       equal length are very often lines of the same shape, whose walks sit in
       the same states at the same time. */
    auto bucketOf = [](uint32_t off, uint32_t len) -> uint32_t {
      return ((off != kUnstaged) && (len != 0)) ? (kSortBuckets - 1) - min(kSortBuckets - 2, len) : kSortBuckets - 1;
    };
    // counting sort: bucket[] in, position in the sorted order out
    auto sortPositions = [&](const uint32_t (&bucket)[kPer], const bool (&valid)[kPer], uint32_t (&pos)[kPer]) {
      static_assert(kSortBuckets == 256, "every lane scans eight buckets");
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(histA + 32 * lane), "r"(0u) : "memory");
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(histA + 32 * lane + 16), "r"(0u) : "memory");
      __syncwarp();
#pragma unroll
      for (uint32_t k = 0; k < kPer; ++k)
        pos[k] = valid[k] ? atomicAdd(sHist + bucket[k], 1u) : 0;
      __syncwarp();
      // exclusive prefix over the buckets: lane l owns buckets 8 l .. 8 l + 7
      const uint4 h0 = lds128(histA + 32 * lane), h1 = lds128(histA + 32 * lane + 16);
      const uint32_t mine = h0.x + h0.y + h0.z + h0.w + h1.x + h1.y + h1.z + h1.w;
      uint32_t incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= static_cast<uint32_t>(d))
          incl += up;
      }
      uint4 e0, e1;
      e0.x = incl - mine;
      e0.y = e0.x + h0.x;
      e0.z = e0.y + h0.y;
      e0.w = e0.z + h0.z;
      e1.x = e0.w + h0.w;
      e1.y = e1.x + h1.x;
      e1.z = e1.y + h1.y;
      e1.w = e1.z + h1.z;
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(histA + 32 * lane), "r"(e0.x), "r"(e0.y), "r"(e0.z), "r"(e0.w) : "memory");
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(histA + 32 * lane + 16), "r"(e1.x), "r"(e1.y), "r"(e1.z), "r"(e1.w) : "memory");
      __syncwarp();
#pragma unroll
      for (uint32_t k = 0; k < kPer; ++k)
        if (valid[k])
          pos[k] += lds32(histA + 4 * bucket[k]);
      __syncwarp();
    };
    auto finishBatch = [&](uint32_t s, uint64_t first, uint32_t count) {
      if (lane == 0) {
        *reinterpret_cast<uint64_t *>(smem + hp.oMeta + s * hp.metaBytes + kMetaFirst) = first;
        sts32(sBase + hp.oMeta + s * hp.metaBytes + kMetaCnt, count);
        sts32(sBase + hp.oMeta + s * hp.metaBytes + kMetaSeq, q);
      }
      __syncwarp();
      if (lane == 0)
        mbar_arrive(barsAddr + 8 * s);
    };

    /* The rare shapes: lines [i0, cnt) of the sub-block at `base`, read from
       global memory.  Ranges that lie back to back and fit a slot go in with
       one bulk copy (a range that does not fit is halved); below a tile's
       worth the lines are gathered one by one (views that are not in text
       order, lines of many kilobytes), one bulk copy per line; a line longer
       than a slot is left to the consumers' global-memory walk. */
    auto slowPath = [&](uint64_t base, uint32_t cnt) {
      auto startOf = [&](uint32_t i) { return textA + __ldg(args.starts + base + i); };
      auto lenOf = [&](uint32_t i) { return __ldg(args.lens + base + i); };
      // lines [i0, i1) with offsets in offA: sorted straight into the slot's arrays
      auto publishRange = [&](uint32_t s, uint32_t i0, uint32_t i1) {
        const uint32_t mA = sBase + hp.oMeta + s * hp.metaBytes;
        uint32_t bucket[kPer], pos[kPer];
        bool valid[kPer];
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k) {
          const uint32_t i = i0 + lane + 32 * k;
          valid[k] = i < i1;
          bucket[k] = valid[k] ? bucketOf(lds16(offA + 2 * i), lenOf(i)) : 0;
        }
        sortPositions(bucket, valid, pos);
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k) {
          const uint32_t i = i0 + lane + 32 * k;
          if (valid[k]) {
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(mA + kMetaIdx + pos[k]), "r"(i - i0) : "memory");
            sts32(mA + kMetaOL + 4 * pos[k], lds16(offA + 2 * i) | (min(lenOf(i), 0xffffu) << 16));
          }
        }
        finishBatch(s, base + i0, i1 - i0);
      };
      uint32_t i0 = 0, tryLines = cnt;
      while (i0 < cnt) {
        const uint32_t i1 = min(i0 + tryLines, cnt);
        const uint64_t a0 = startOf(i0);
        const uint64_t aEnd = startOf(i1 - 1) + lenOf(i1 - 1);
        const uint64_t loA = a0 & ~15ull;
        const uint64_t hiA = (aEnd + 15) & ~15ull;
        bool inside = (aEnd >= a0) && (hiA - loA <= hp.slotBytes);
        if (inside) {
          bool mine = true;
          for (uint32_t i = i0 + lane; i < i1; i += 32) {
            const uint64_t a = startOf(i);
            mine &= (a >= a0) && (a + lenOf(i) <= aEnd);
            sts16(offA + 2 * i, static_cast<uint32_t>(a - loA));
          }
          inside = __all_sync(kFullMask, mine);
        }
        if (inside) {
          const uint32_t s = acquire();
          const uint32_t bytes = static_cast<uint32_t>(hiA - loA);
          if ((lane == 0) && bytes) {
            mbar_expect_tx(barsAddr + 8 * s, bytes);
            bulk_g2s(sBase + hp.oSlots + s * hp.slotBytes, reinterpret_cast<const void *>(loA), bytes,
                     barsAddr + 8 * s);
          }
          __syncwarp();
          publishRange(s, i0, i1);
          i0 = i1;
          continue;
        }
        if (tryLines > 32) {
          tryLines = (tryLines + 1) >> 1;
          continue;
        }
        uint32_t running = 0, g1 = i0;
        bool cut = false;
        for (uint32_t g = i0; (g < cnt) && !cut; g += 32) {
          const uint32_t i = g + lane;
          const bool valid = i < cnt;
          uint32_t ext = 0, low = 0;
          bool big = false;
          if (valid) {
            const uint64_t a = startOf(i);
            const uint64_t len = lenOf(i);
            low = static_cast<uint32_t>(a & 15);
            const uint64_t e = len ? ((a + len + 15) & ~15ull) - (a & ~15ull) : 0;
            big = e > hp.slotBytes;
            ext = big ? 0 : static_cast<uint32_t>(e);
          }
          uint32_t incl = ext;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= static_cast<uint32_t>(d))
              incl += up;
          }
          const bool fits = !valid || (running + incl <= hp.slotBytes);
          const unsigned bad = ~__ballot_sync(kFullMask, fits);
          const uint32_t nfit = bad ? static_cast<uint32_t>(__ffs(bad) - 1) : 32u;
          if (valid && (lane < nfit))
            sts16(offA + 2 * i, big ? kUnstaged : running + incl - ext + low);
          g1 = min(g + nfit, cnt);
          if (g1 < min(g + 32, cnt))
            cut = true;
          else
            running += __shfl_sync(kFullMask, incl, 31);
        }
        __syncwarp();
        const uint32_t s = acquire();
        const uint32_t slotA = sBase + hp.oSlots + s * hp.slotBytes;
        uint32_t mineBytes = 0;
        for (uint32_t i = i0 + lane; i < g1; i += 32) {
          const uint32_t off = lds16(offA + 2 * i);
          const uint64_t len = lenOf(i);
          if ((off != kUnstaged) && len) {
            const uint64_t a = startOf(i);
            const uint32_t ext = static_cast<uint32_t>(((a + len + 15) & ~15ull) - (a & ~15ull));
            bulk_g2s(slotA + (off & ~15u), reinterpret_cast<const void *>(a & ~15ull), ext, barsAddr + 8 * s);
            mineBytes += ext;
          }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1)
          mineBytes += __shfl_xor_sync(kFullMask, mineBytes, d);
        if ((lane == 0) && mineBytes)
          mbar_expect_tx(barsAddr + 8 * s, mineBytes);
        publishRange(s, i0, g1);
        i0 = g1;
      }
    };

    unsigned long long claimed = 0;
    if (lane == 0)
      claimed = atomicAdd(blockCounter, 1ull);
    uint64_t block = __shfl_sync(kFullMask, claimed, 0);
    const long long tStart = dbgOn ? clock64() : 0;
    /* The index entries of a sub-block (lane + 32 k: coalesced, all in flight
       at once) are loaded one sub-block ahead - as soon as the registers of the
       current one are free - so that their latency lies behind the wait for a
       slot and the hand-over instead of in front of the sort. */
    uint64_t st[kPer];
    uint32_t ln[kPer];
    bool valid[kPer];
    auto loadIndex = [&](uint64_t blk) {
      const uint64_t base = blk * SUB;
      const uint32_t cnt = (base < n) ? static_cast<uint32_t>(min(static_cast<uint64_t>(SUB), n - base)) : 0;
#pragma unroll
      for (uint32_t k = 0; k < kPer; ++k) {
        const uint32_t i = lane + 32 * k;
        valid[k] = i < cnt;
        st[k] = valid[k] ? __ldg(args.starts + base + i) : 0;
        ln[k] = valid[k] ? __ldg(args.lens + base + i) : 0;
      }
    };
    loadIndex(block);
    while (block * SUB < n) {
      const long long tI0 = dbgOn ? clock64() : 0;
      const uint64_t base = block * SUB;
      const uint32_t cnt = static_cast<uint32_t>(min(static_cast<uint64_t>(SUB), n - base));
      if (lane == 0)
        claimed = atomicAdd(blockCounter, 1ull);  // the next sub-block; looked at below
      /* The span of the sub-block's text, in 32 bits: every line's start
         relative to the sub-block's first line (lane 0 has it; a sub-block that
         is taken has one), smallest start and largest end by REDUX.  (In 64
         bits it was eight wide adds, mins and maxes per lane and five rounds of
         four shuffles.)  Lines a gigabyte apart go the rare shapes' way. */
      const uint64_t ref = textA + __shfl_sync(kFullMask, st[0], 0);
      int32_t rel[kPer];
      int32_t dlo = 0x7fffffff, dhi = -0x7fffffff - 1;
      bool near = true;
#pragma unroll
      for (uint32_t k = 0; k < kPer; ++k) {
        rel[k] = 0;
        if (valid[k]) {
          const int64_t d = static_cast<int64_t>(st[k] + textA - ref);
          near = near && (d > -(1ll << 30)) && (d < (1ll << 30)) && (ln[k] < (1u << 30));
          rel[k] = static_cast<int32_t>(d);
          dlo = min(dlo, rel[k]);
          dhi = max(dhi, rel[k] + static_cast<int32_t>(ln[k]));
        }
      }
      if (dbgOn)
        tIdx += clock64() - tI0;
      near = __all_sync(kFullMask, near);
      uint64_t lo = ~0ull, hi = 0;  // (hi < lo: not the usual shape)
      if (near) {
        dlo = __reduce_min_sync(kFullMask, dlo);
        dhi = __reduce_max_sync(kFullMask, dhi);
        lo = ref + static_cast<int64_t>(dlo);
        hi = ref + static_cast<int64_t>(dhi);
      }
      const uint64_t loA = lo & ~15ull;
      const uint64_t hiA = (hi + 15) & ~15ull;
      if ((hi >= lo) && (hiA - loA <= hp.slotBytes)) {
        // the usual shape: the whole sub-block's text fits a slot (whatever the
        // order of its lines).  Sort in the scratch, then take a slot, start
        // the copy and hand the sorted arrays over.  The text is asked into L2
        // now: the copy itself starts only when a slot is free, and what it
        // takes from there on is time the slot stands empty.
        if ((lane == 0) && (hiA > loA))
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(loA), "r"(static_cast<uint32_t>(hiA - loA)) : "memory");
        const uint32_t relBias = static_cast<uint32_t>(ref - loA);  // a line's offset in the slot = rel + (ref - loA), mod 2^32
        uint32_t bucket[kPer], pos[kPer];
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k)
          bucket[k] = valid[k] ? bucketOf(static_cast<uint32_t>(rel[k]) + relBias, ln[k]) : 0;
        sortPositions(bucket, valid, pos);
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k)
          if (valid[k]) {
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(scrA + kMetaIdx + pos[k]), "r"(lane + 32 * k) : "memory");
            sts32(scrA + kMetaOL + 4 * pos[k], (static_cast<uint32_t>(rel[k]) + relBias) | (min(ln[k], 0xffffu) << 16));
          }
        __syncwarp();
        block = __shfl_sync(kFullMask, claimed, 0);
        loadIndex(block);
        const uint32_t s = acquire();
        const uint32_t bytes = static_cast<uint32_t>(hiA - loA);
        if ((lane == 0) && bytes) {
          mbar_expect_tx(barsAddr + 8 * s, bytes);
          bulk_g2s(sBase + hp.oSlots + s * hp.slotBytes, reinterpret_cast<const void *>(loA), bytes,
                   barsAddr + 8 * s);
        }
        const uint32_t mA = sBase + hp.oMeta + s * hp.metaBytes;
        static_assert((kMetaArrays % 16 == 0) && (kMetaArrays <= 3 * 32 * 16), "the scratch mirrors the slot's arrays, three 16-byte pieces per lane at most");
#pragma unroll
        for (uint32_t k = 0; k < 3; ++k) {
          const uint32_t o = 16 * (lane + 32 * k);
          if (o < kMetaArrays) {
            const uint4 v = lds128(scrA + o);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(mA + o), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        }
        finishBatch(s, base, cnt);
      }
      else {
        slowPath(base, cnt);
        block = __shfl_sync(kFullMask, claimed, 0);
        loadIndex(block);
      }
    }
    if (dbgOn && (lane == 0)) {
      atomicAdd(hp.dbg + 0, static_cast<unsigned long long>(clock64() - tStart));
      atomicAdd(hp.dbg + 1, tIdx);
      atomicAdd(hp.dbg + 2, tAcq);
      atomicAdd(hp.dbg + 3, nBatches);
    }
    // the last producer to run out of lines tells the consumers how many
    // batches there were: a ticket for a later one means the work is done
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(reinterpret_cast<uint32_t *>(smem + hp.oCtl + 8), 1u) == P - 1) {
        __threadfence_block();
        const uint32_t total = lds32_volatile(ctlAddr + 4);
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ctlAddr + 12), "r"(total) : "memory");
        __threadfence_block();
        for (uint32_t s = 0; s < R; ++s)  // wake the sleepers
          mbar_arrive(barsAddr + 8 * (2 * R + s));
      }
    }
    return;
  }

  // ========================= consumers =======================================
  HotCtx hx;
  hx.tab = sBase;
  hx.c2v = sBase + hp.oC2v;
  hx.sink = hd.nHot;
  hx.cLimit = (STYLE == STY_FULL) ? hd.cAbs : hd.cAccept;
  hx.cAbs = hd.cAbs;
  hx.cPure = hd.cPure;
  hx.wide = hd.wideTable;
  hx.codeOf = hd.codeOf;
  unsigned long long tWait = 0, tWalk = 0, tTail = 0, nTiles = 0;
  const long long tStartC = dbgOn ? clock64() : 0;

  /* A tile's ticket is taken while the tile before it is finished off (the
     atomic's round trip lies behind the flagged lanes and the outcomes).  PTX
     with a per-lane addend: atomicAdd() under "lane == 0" becomes the
     warp-aggregated form, whose shuffle waits for the result on the spot. */
  auto takeTicket = [&]() -> uint32_t {
    uint32_t t = 0;
    if (lane == 0) {
      uint32_t l0;
      asm volatile("mov.u32 %0, %%laneid;" : "=r"(l0));
      asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(t) : "r"(ctlAddr), "r"(l0 + 1u) : "memory");
    }
    return t;
  };
  uint32_t tkAhead = takeTicket();
  for (;;) {
    const long long tc0 = dbgOn ? clock64() : 0;
    const uint32_t tk = __shfl_sync(kFullMask, tkAhead, 0);
    // (by multiplication: two divisions by run-time values were 40 instructions
    // and two MUFU round trips in front of every tile)
    const uint32_t qq = (TPB == 1) ? tk : __umulhi(tk, hp.tpbMagic);  // (d = 1 has no such magic in 32 bits)
    const uint32_t t = tk - qq * TPB;
    const uint32_t lap = __umulhi(qq, hp.slotsMagic);  // qq / R
    const uint32_t s = qq - lap * R;
    /* Wait for batch qq.  A parity wait tells two consecutive phases of
       full[s] apart and no more: with several producers batches are published
       out of order, and this warp may be here before batch qq - R - same slot -
       is even published.  The slot's sequence number says whose phase it was. */
    bool ended = false;
    const uint32_t seqAddr = sBase + hp.oMeta + s * hp.metaBytes + kMetaSeq;
    /* The bytes of a bulk copy arrive at full[s] piece by piece, and every
       piece wakes whoever waits there (measured: ~50 wake-ups per wait, the
       waiters' instructions a quarter of all the kernel issued).  So only the
       warp with the batch's first tile waits at full[s]; it then arrives at
       ready[s], where the others sleep through the copy.  ready[s] is a place
       to sleep and no more - the last producer arrives there too, to wake
       those whose batch will never come: whether the batch is in is always
       read off full[s] and the slot's sequence number. */
    const uint32_t par = lap & 1;
    const uint32_t fullBar = barsAddr + 8 * s, readyBar = barsAddr + 8 * (2 * R + s);
    for (uint32_t spins = 0;; ++spins) {
      const bool woke = (t != 0) && mbar_try_wait(readyBar, par);
      const bool landed = (t != 0) ? mbar_test_wait(fullBar, par) : mbar_try_wait(fullBar, par);
      if (landed) {
        if (lds32_volatile(seqAddr) == qq)
          break;
        __nanosleep(200);  // an earlier phase: batch qq - R is still to come
      }
      // not there yet: is there such a batch at all?  (a poll is a
      // shared-memory wavefront the walkers want: the sleepers look when they
      // are woken, and every so often)
      if ((t == 0) || woke || ((spins & 7) == 7)) {
        if (qq >= lds32_volatile(ctlAddr + 12)) {
          ended = true;
          break;
        }
      }
    }
    if (ended)
      break;
    if ((t == 0) && (lane == 0))
      mbar_arrive(barsAddr + 8 * (2 * R + s));
    const long long tc1 = dbgOn ? clock64() : 0;
    const uint32_t mA = sBase + hp.oMeta + s * hp.metaBytes;
    // (the entries are read whether the batch has them or not - pos < 256, the
    // arrays are that long - so that the four loads are under way together)
    const uint32_t pos = t * 32 + lane;
    uint32_t idx = lds8(mA + kMetaIdx + pos);
    const uint32_t ol = lds32(mA + kMetaOL + 4 * pos);
    uint32_t off = ol & 0xffffu;
    uint32_t len = ol >> 16;
    const uint32_t cnt = lds32(mA + kMetaCnt);
    bool has = pos < cnt;
    if (!has) {
      idx = 0;
      off = kUnstaged;
      len = 0;
    }
    /* Deal the tile's lines out to the lanes by the 16-byte bank group their
       text starts in (every chunk of a line is then in step with its first):
       an LDS.128 is served eight lanes at a time, and eight lanes that read
       from eight different groups cost one wavefront (measured with the lines
       as they come: 10 wavefronts per instruction; 4 is the floor).  The
       lines are ranked by group; rank r goes to lane (r % 4) * 8 + r / 4, so
       every run of eight lanes gets every fourth line of that order. */
    {
      const uint32_t g = (off >> 4) & 7;  // (no line: 7, sorts last)
      /* Rank by (group, lane).  How many lanes each group has: a byte per
         group, four groups to a word, summed over the warp (REDUX; a count is
         at most 32); the lanes of the groups below g are the bytes below g's,
         added up by a multiplication.  (Eight ballots with their counts were
         4 % of the kernel's stall samples.) */
      const uint32_t sh = (g & 3u) * 8u;
      const uint32_t one = 1u << sh;
      const uint32_t tlo = __reduce_add_sync(kFullMask, (g < 4) ? one : 0u);
      const uint32_t thi = __reduce_add_sync(kFullMask, (g < 4) ? 0u : one);
      const uint32_t below = one - 1u;
      const uint32_t part = (((g < 4) ? tlo : thi) & below) * 0x01010101u;
      const uint32_t whole = (g < 4) ? 0u : tlo * 0x01010101u;
      const unsigned same = __match_any_sync(kFullMask, g);
      const uint32_t r = ((part + whole) >> 24) + __popc(same & ((1u << lane) - 1));
      const uint32_t to = ((r & 3) << 3) | (r >> 2);
      const uint32_t dealA = sBase + hp.oDeal + 32 * warp;
      asm volatile("st.shared.u8 [%0], %1;" ::"r"(dealA + to), "r"(lane) : "memory");
      __syncwarp();
      const uint32_t src = lds8(dealA + lane);
      __syncwarp();
      idx = __shfl_sync(kFullMask, idx, src);
      off = __shfl_sync(kFullMask, off, src);
      len = __shfl_sync(kFullMask, len, src);
      has = __shfl_sync(kFullMask, has ? 1u : 0u, src) != 0;
    }
    // what the lane keeps of its line while it walks: where it was dealt from
    // (idx, off and len are fetched again when the walk is over)
    const uint32_t slotA = sBase + hp.oSlots + s * hp.slotBytes;
    Walk<STYLE, false> wk;
    uint32_t code = hd.initCode;
    bool done = false;  // the walk stopped, or it went on the full table
    /* What the outcome needs of wk, in registers of their own: wk goes through
       calls by reference (careful passes, the full table) and so lives in local
       memory, and a tile's lanes nearly never take those paths - but reading
       the two fields back from there was a round trip to L2 for every tile
       (ncu: 2 % of all the kernel's stall samples on that one load). */
    uint32_t oMend = 0;
    bool oAbsorbed = false;
    /* The slot is given back as soon as the tile's lanes need nothing of it any
       more - behind the walk, in front of the flagged lanes' second look (which
       reads no text) and the outcomes - unless a lane has to go byte by byte
       through its text: the producers get the slot a few hundred cycles
       earlier for every tile, and what the consumers wait for is the copy that
       follows.  The batch's first line number is read before that (the slot's
       meta block is the next batch's then). */
    uint64_t first = 0;
    bool released = false;
    auto release = [&]() {
      first = *reinterpret_cast<const uint64_t *>(smem + hp.oMeta + s * hp.metaBytes + kMetaFirst);
      __syncwarp();
      if (lane == 0)
        mbar_arrive(barsAddr + 8 * (R + s));
      released = true;
    };

    bool active = has && (off != kUnstaged) && (len != 0);
    const uint32_t tb = slotA + off;  // the line's first byte
    Feed fd;
    fd.blk = 0;
    fd.cur = make_uint4(0, 0, 0, 0);
    fd.sel = 0x3210u;
    fd.w1 = fd.w2 = false;
    if (active)
      feed_begin(fd, tb);
    uint32_t rem = active ? len : 0;  // bytes of the line from the current chunk on
    // chunks that are whole for every lane of the tile: a loop without masks
    const uint32_t kFull = __reduce_min_sync(kFullMask, active ? (len >> 4) : 0xffffffffu);

    if (DEFER) {
      /* Lean walk for the styles that end at the first state that needs a
         decision, on an image where such a state sticks.  Live per lane: the
         text feed, the bytes left, the state in front of the chunk; a flagged
         lane freezes them and steps aside. */
      uint32_t fc = 0xffffffffu;  // flagged: the code the chunk's run ended in
      /* A flagged lane keeps the state in front of the QUARTER of the chunk in
         which its run met the deciding code (such codes stick, so the codes
         behind the quarters tell which it was): rem and code move up to that
         quarter, and the second look is at four bytes, not sixteen. */
      uint32_t fw = 0;  // the four bytes of that quarter (the second look needs no text from shared memory)
      auto flag = [&](uint32_t c2, uint32_t cp0, uint32_t cp1, uint32_t cp2, const uint4 &w) {
        fc = c2;
        active = false;
        fw = w.x;
        if (c2 != 255u) {  // (255: possibly a byte >= 128 on 128-column rows - the codes mean nothing)
          const uint32_t lim = hx.cLimit;
          if (cp0 < lim) {
            uint32_t adv = 4, cq = cp0;
            fw = w.y;
            if (cp1 < lim) {
              adv = 8;
              cq = cp1;
              fw = w.z;
              if (cp2 < lim) {
                adv = 12;
                cq = cp2;
                fw = w.w;
              }
            }
            code = cq;
            rem -= adv;
          }
        }
      };
      if (kFull != 0xffffffffu) {
#pragma unroll kHotUnroll
        for (uint32_t k = 0; k < kFull; ++k) {
          if (active) {
            const uint4 w = feed_next(fd);
            uint32_t cp0, cp1, cp2;
            const uint32_t c2 = hot_run_q<true, STRIDE>(hx, w, 16, code, cp0, cp1, cp2);
            if (c2 >= hx.cLimit)
              flag(c2, cp0, cp1, cp2, w);
            else {
              code = c2;
              rem -= 16;
            }
          }
        }
      }
      active = active && (rem != 0);
      while (__any_sync(kFullMask, active)) {
        if (active) {
          const uint4 w = feed_next(fd);
          uint32_t cp0, cp1, cp2;
          const uint32_t c2 = hot_run_q<false, STRIDE>(hx, w, min(16u, rem), code, cp0, cp1, cp2);
          if (c2 >= hx.cLimit)
            flag(c2, cp0, cp1, cp2, w);
          else {
            code = c2;
            if (rem <= 16)
              active = false;
            else
              rem -= 16;
          }
        }
      }
      tkAhead = takeTicket();
      if (!__any_sync(kFullMask, (fc != 0xffffffffu) && ((fc == 255u) || (fc == hx.sink))))
        release();
      if (__any_sync(kFullMask, fc != 0xffffffffu)) {
        /* Flagged lanes, together.  Nearly always the run ended in a state
           that settles the line: an accepting one (Instant: where it was
           entered is counted out by a second run over the quarter), an
           absorbing or a dead one (no position needed).  What is left - the
           walk left the hot set, a byte >= 128 on 128-column rows - goes byte
           by byte from where the lane stopped, to the line's end if it has to. */
        const bool settled = (fc != 0xffffffffu) && (fc != 255u) && (fc != hx.sink);
        if (settled) {
          if ((STYLE == STY_INSTANT) && (fc < hx.cPure)) {
            oMend = (len - rem) + hot_locate_quarter<STRIDE>(hx, fw, min(4u, rem), code) + 1;
            done = true;
          }
          code = fc;
        }
        else if (fc != 0xffffffffu) {
          wk.begin(pd);
          hot_careful<STYLE, STRIDE>(pd, hx, tb, len - rem, len, 0, code, wk);
          oMend = wk.mend;
          oAbsorbed = wk.absorbed;
          done = true;
        }
      }
    }
    else {
      wk.begin(pd);
      done = true;
      uint32_t at = 0;  // position of the current chunk in the line
      if (kFull != 0xffffffffu) {
        for (uint32_t k = 0; k < kFull; ++k) {
          if (active) {
            const uint4 w = feed_next(fd);
            if (hot_chunk<STYLE, true, STRIDE>(pd, hx, w, tb + at, 0, 16, at, code, wk))
              active = false;
            at += 16;
          }
        }
      }
      while (__any_sync(kFullMask, active)) {
        if (active) {
          if (at >= len)
            active = false;
          else {
            const uint4 w = feed_next(fd);
            if (hot_chunk<STYLE, false, STRIDE>(pd, hx, w, tb + at, 0, min(16u, len - at), at, code, wk))
              active = false;
            at += 16;
          }
        }
      }
      oMend = wk.mend;
      oAbsorbed = wk.absorbed;
      tkAhead = takeTicket();
      release();
    }
    if (!released)
      first = *reinterpret_cast<const uint64_t *>(smem + hp.oMeta + s * hp.metaBytes + kMetaFirst);

    if (has) {
      uint32_t lineLen = len;
      if (off == kUnstaged) {
        // longer than a slot: from global memory, on the full table
        const uint64_t line = first + idx;
        lineLen = __ldg(args.lens + line);
        wk.begin(pd);
        done = true;
        cold_bytes<STYLE>(pd, hx.wide, wk, args.text + __ldg(args.starts + line), lineLen, 0);
        oMend = wk.mend;
        oAbsorbed = wk.absorbed;
        code = hx.sink;  // the state is in wk.v
      }
      // Walk::finish (Matcher.h:481-493), with the result ids of hot states
      // taken from shared memory
      int32_t now, before = 0;
      if (code != hx.sink)
        now = static_cast<int32_t>(lds32(hx.c2v + 1024 + 4 * code));
      else
        now = ((wk.v >= pd.acceptV) && (wk.v < pd.pureV)) ? __ldg(pd.results + state_index(pd, wk.v)) : 0;
      if ((STYLE == STY_LAST) && (wk.accV != kNone))
        before = (wk.accV & 0x80000000u) ? static_cast<int32_t>(wk.accV & 0x7fffffffu)
                                          : __ldg(pd.results + state_index(pd, wk.accV));
      const int32_t result = (before > 0) ? before : now;
      // (a walk that ran out without a decision has no accepting position to report)
      const bool absorbed = done && oAbsorbed;
      const uint32_t mend = done ? oMend : 0;
      const uint32_t end = (result > 0) ? (((STYLE == STY_FULL) || absorbed) ? lineLen : mend) : 0;
      // straight to the line's place: the 256 outcomes of a batch are 1 KB per
      // array, written within a few microseconds of each other - L2 puts the
      // sectors together (collecting them per batch in shared memory for a
      // coalesced copy cost a counter, two fences and 570 cycles per tile)
      const uint64_t line = first + idx;
      args.results[line] = result;
      if (args.matchEnds)
        args.matchEnds[line] = end;
    }
    __syncwarp();
    const long long tc2 = dbgOn ? clock64() : 0;
    if (!released && (lane == 0))
      mbar_arrive(barsAddr + 8 * (R + s));
    if (dbgOn) {
      const long long tc3 = clock64();
      tWait += tc1 - tc0;
      tWalk += tc2 - tc1;
      tTail += tc3 - tc2;
      ++nTiles;
    }
  }
  if (dbgOn && (lane == 0)) {
    atomicAdd(hp.dbg + 4, static_cast<unsigned long long>(clock64() - tStartC));
    atomicAdd(hp.dbg + 5, tWait);
    atomicAdd(hp.dbg + 6, tWalk);
    atomicAdd(hp.dbg + 7, tTail);
    atomicAdd(hp.dbg + 8, nTiles);
  }
}

// ---- profile ----------------------------------------------------------------

__global__ void __launch_bounds__(256)
profile_kernel(const ProgramDev pd, const uint32_t wide, const BatchArgs args, const uint64_t stride,
               const uint32_t nStates, unsigned long long *hist, unsigned long long *lenSum) {
  const AnyTable tab = any_table(pd, wide);
  const uint64_t n = args.nLinesDev ? min(*args.nLinesDev, args.nLines) : args.nLines;
  const uint64_t samples = (n + stride - 1) / stride;
  unsigned long long bytes = 0, highBytes = 0;
  for (uint64_t sIdx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; sIdx < samples;
       sIdx += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t line = sIdx * stride;
    const uint8_t *p = args.text + __ldg(args.starts + line);
    const uint32_t len = __ldg(args.lens + line);
    bytes += len;
    uint32_t v = pd.initV;
    unsigned long long *h = hist;  // until the first accepting state: what an Instant walk enters
    uint32_t initRuns = 1;         // entries into the initial state, added once per line
    for (uint32_t i = 0; i < len; ++i) {
      const uint32_t byte = __ldg(p + i);
      highBytes += byte >> 7;
      const uint32_t nv = tab.next(v, __ldg(pd.clsOffset + byte));
      if (nv != v) {
        if ((nv == pd.initV) && (h == hist))
          ++initRuns;
        else
          atomicAdd(h + state_index(pd, nv), 1ull);
        v = nv;
        if (v >= pd.acceptV)
          h = hist + nStates;
        if (v >= pd.pureV)
          break;
      }
    }
    atomicAdd(hist + state_index(pd, pd.initV), static_cast<unsigned long long>(initRuns));
  }
  if (bytes)
    atomicAdd(lenSum, bytes);
  if (highBytes)
    atomicAdd(lenSum + 1, highBytes);
}

} // anonymous

cudaError_t launch_profile(const ProgramDev &pd, uint32_t wideTable, const BatchArgs &args,
                           uint64_t stride, uint32_t nStates, unsigned long long *hist,
                           unsigned long long *lenSum, cudaStream_t stream) {
  if ((args.nLines == 0) || (stride == 0))
    return cudaSuccess;
  const uint64_t samples = (args.nLines + stride - 1) / stride;
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>((samples + 255) / 256, 148 * 8));
  profile_kernel<<<grid, 256, 0, stream>>>(pd, wideTable, args, stride, nStates, hist, lenSum);
  return cudaGetLastError();
}

cudaError_t launch_batch_hot(const ProgramDev &pd, const HotDev &hd, const BatchArgs &args,
                             int style, const LaunchGeometry &geo, const HotTuning &tune,
                             uint32_t avgLineBytes, unsigned long long *counter,
                             cudaStream_t stream, int *launches) {
  if ((style != STY_INSTANT) && (style != STY_LAST) && (style != STY_FULL))
    return cudaErrorNotSupported;
  if (args.checkLeader || args.matchStarts || !hd.table)
    return cudaErrorNotSupported;
  // A lane that leaves the hot set walks on the full table in global memory,
  // a byte per L2 round trip, with its warp waiting: with more than the odd
  // excursion the other kernels are faster (measured on union64, styles Last
  // and Full, where a line's walk goes on behind its match and enters more
  // states than the hot set holds: 560-700 GB/s against 1500-1650).
  if ((hd.coverage < 999) && !tune.force)
    return cudaErrorNotSupported;
  if (args.nLines == 0)
    return cudaSuccess;

  HotParams hp{};
  uint32_t at = (hd.tableBytes + 15u) & ~15u;
  auto take = [&](uint32_t bytes, uint32_t align) {
    at = (at + align - 1) & ~(align - 1);
    const uint32_t o = at;
    at += bytes;
    return o;
  };
  const uint32_t producers = std::max(1, std::min(tune.producers ? tune.producers : 3, static_cast<int>(kMaxProducers)));
  hp.nProducers = producers;
  hp.prodBytes = (kProdBytes + 15u) & ~15u;
  hp.oC2v = take(2048, 16);
  hp.oBars = take(3 * kMaxSlots * 8, 16);
  hp.oCtl = take(32, 16);
  hp.oProd = take(producers * hp.prodBytes, 16);
  hp.oDeal = take(32 * 32, 16);
  hp.metaBytes = (kMetaBytes + 15u) & ~15u;

  // Geometry.  A batch's text has to fit a slot; the ring has to hold a tile
  // for every consumer warp plus one batch that is being fetched
  // (consumers <= (slots - 1) * tiles per batch: that also keeps a ticket from
  // ever being two phases away from its slot's barrier).
  const int maxConsumers = 32 - static_cast<int>(producers);
  uint32_t batchLines = tune.batchLines ? static_cast<uint32_t>(tune.batchLines) : kMaxBatch;
  batchLines = std::max(32u, std::min(batchLines, kMaxBatch)) & ~31u;
  const uint32_t fixed = at + 256 + 256;  // alignment pad in front, slack for the slots' own alignment
  const uint32_t room = (static_cast<uint32_t>(geo.maxSmemOptin) > fixed) ? geo.maxSmemOptin - fixed : 0;
  uint32_t slots = 0, slotBytes = 0;
  int warps = 0;
  for (;;) {
    const uint32_t tiles = batchLines / 32;
    // a batch's text: its mean plus 1.5 sigma if line lengths scatter by a
    // quarter of their mean (a batch that does not fit is split, see the producer)
    const uint32_t sqrtB = (batchLines >= 256) ? 16 : (batchLines >= 128) ? 12 : (batchLines >= 64) ? 8 : 6;
    const uint32_t need = avgLineBytes ? (avgLineBytes + 2) * batchLines + avgLineBytes * sqrtB * 3 / 8 + 64
                                       : 108u * batchLines;
    slotBytes = std::max(256u, std::min((need + 15u) & ~15u, 65520u));
    if (tune.slotBytes)
      slotBytes = std::max(256u, std::min((static_cast<uint32_t>(tune.slotBytes) + 15u) & ~15u, 65520u));
    slots = std::min<uint32_t>(kMaxSlots, room / (slotBytes + hp.metaBytes));
    if (tune.slots)
      slots = std::min<uint32_t>(slots, static_cast<uint32_t>(tune.slots));
    if (slots >= 2) {
      if (!tune.slotBytes)  // the room that is left makes the slots longer
        slotBytes = std::min((room / slots - hp.metaBytes) & ~15u, 65520u);
      warps = tune.consumerWarps ? tune.consumerWarps : maxConsumers;
      warps = std::max(1, std::min(warps, std::min(maxConsumers, static_cast<int>((slots - 1) * tiles))));
      if ((warps >= 16) || (batchLines == 32) || tune.batchLines)
        break;
    }
    if (batchLines == 32)
      return cudaErrorNotSupported;
    batchLines -= 32;
  }
  hp.dbg = tune.debug;
  hp.nSlots = slots;
  hp.slotsMagic = static_cast<uint32_t>((1ull << 32) / slots) + 1u;
  hp.tpbMagic = static_cast<uint32_t>((1ull << 32) / (batchLines / 32)) + 1u;
  hp.slotBytes = slotBytes;
  hp.batchLines = batchLines;
  hp.tilesPerBatch = batchLines / 32;
  hp.oMeta = take(slots * hp.metaBytes, 16);
  hp.oSlots = take(slots * slotBytes, 128);
  const size_t smemBytes = at + 256 + 64;  // alignment pad; a feed reads up to two blocks (and a word) past its line's last chunk
  if (smemBytes > static_cast<size_t>(geo.maxSmemOptin))
    return cudaErrorNotSupported;

  cudaError_t err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
  if (err != cudaSuccess)
    return err;
  // careful passes together at the end of a tile - unless the hot set leaves
  // so much out that lanes would often go on alone behind theirs
  bool defer = tune.defer ? (tune.defer > 0) : (hd.coverage >= 999);
  if ((style == STY_INSTANT) && !hd.acceptSticks)
    defer = false;  // the lean walk needs an image whose accepting codes stick
  constexpr int W = kHotStrideWide, A = kHotStrideAscii;
  const bool ascii = hd.stride == kHotStrideAscii;
  const bool dbg = tune.debug != nullptr;
#define HOT_K(S, D, ST) (dbg ? batch_hot_kernel<S, D, ST, true> : batch_hot_kernel<S, D, ST, false>)
  auto kern = (style == STY_LAST)
      ? (ascii ? HOT_K(STY_LAST, false, A) : HOT_K(STY_LAST, false, W))
      : (style == STY_INSTANT)
          ? (defer ? (ascii ? HOT_K(STY_INSTANT, true, A) : HOT_K(STY_INSTANT, true, W))
                   : (ascii ? HOT_K(STY_INSTANT, false, A) : HOT_K(STY_INSTANT, false, W)))
          : (defer ? (ascii ? HOT_K(STY_FULL, true, A) : HOT_K(STY_FULL, true, W))
                   : (ascii ? HOT_K(STY_FULL, false, A) : HOT_K(STY_FULL, false, W)));
#undef HOT_K
  err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemBytes));
  if (err != cudaSuccess)
    return err;
  const uint64_t blocks = (args.nLines + batchLines - 1) / batchLines;
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(geo.smCount, blocks));
  kern<<<grid, (warps + static_cast<int>(producers)) * 32, smemBytes, stream>>>(pd, hd, args, hp, counter);
  err = cudaGetLastError();
  if ((err == cudaSuccess) && launches)
    *launches += 1;
  return err;
}

} // namespace red_b200
