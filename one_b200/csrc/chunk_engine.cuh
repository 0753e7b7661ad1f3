/* chunk_engine.cuh - the chunk engine shared by the batch kernels
   (kernels_lanes.cu) and the long-stream matcher (stream_match.cu): fast-pass
   policies per table layout and process_chunk, which walks one aligned
   16-byte chunk per lane.  See kernels_lanes.cu for the description. */

#pragma once

#include "device_program.h"
#include "walk.cuh"

#include <type_traits>

namespace red_b200 {
namespace dev {

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
              Walk<STYLE, WANT_START> &wk, const bool cValid, const bool allFull = false,
              const uint32_t quarters = 0xfu) {
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
        if (allFull) {
          // every lane of the warp that is in this chunk owns all 16 bytes
          // (interior chunks of length-sorted tiles): no per-byte selects
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t word = word_of(cur, q);
            uint32_t m = 0;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint32_t ne = fast.step(e, word, jj);
              if (WANT_START) {
                if ((e == eInit) && (ne != eInit))
                  msFast = pos0 + 4 * q + jj;
              }
              e = ne;
              m = max(m, e);
            }
            if (q == 0) { mq0 = m; cp0 = e; }
            if (q == 1) { mq1 = m; cp1 = e; }
            if (q == 2) { mq2 = m; cp2 = e; }
            if (q == 3) { mq3 = m; }
          }
        }
        else {
          // `quarters` (warp-uniform): bit q clear = no lane of the warp has a
          // byte in quarter q of this chunk (tiles sorted by the alignment of
          // their lines' ends): the quarter's eight lookups are not issued
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t word = word_of(cur, q);
            uint32_t m = 0;
            if (quarters & (1u << q)) {
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
            }
            if (q == 0) { mq0 = m; cp0 = e; }
            if (q == 1) { mq1 = m; cp1 = e; }
            if (q == 2) { mq2 = m; cp2 = e; }
            if (q == 3) { mq3 = m; }
          }
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
                      // without start tracking mstart is free: it reports how
                      // many bytes the walk consumed (kernels_sorted.cu, phase 2)
                      if (!WANT_START && stopped)
                        wk.mstart = pos0 + 4 * q + jj + 1;
                    }
                  }
                }
              }
              else {
                wk.v = fast.dec(eq);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  if (!stopped && (nib & (1u << jj))) {
                    stopped = wk.step(pd, tab, cls[byte_of(word, jj)], pos0 + 4 * q + jj);
                    if (!WANT_START && stopped)
                      wk.mstart = pos0 + 4 * q + jj + 1;
                  }
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

/* Picks the fast-pass policy for a table layout and binds it to the staged
   shared memory (class table at smem, transition table behind it). */
template <bool SMEM, int K, class ClsT> struct Engine {
  using Tab = typename std::conditional<SMEM, SmemTable<K>, GlobalTable>::type;
  static constexpr bool kPtx = SMEM && (sizeof(ClsT) == 1);
  static constexpr bool kByte = SMEM && (K == 9);
  using Fast = typename std::conditional<kByte, SmemFastByte,
      typename std::conditional<kPtx, SmemFast8<K>, GenericFast<Tab, ClsT>>::type>::type;

  // smem: 256-byte aligned start of the staged tables
  static __device__ __forceinline__ Tab table(const ProgramDev &pd, const unsigned char *smem) {
    if constexpr (SMEM) {
      return Tab{smem + kClsBytes};
    }
    else {
      Tab t{static_cast<const unsigned char *>(pd.table)};
      t.hotAddr = smem_u32(smem + kClsBytes);
      t.hotBytes = pd.hotBytes;
      return t;
    }
  }
  static __device__ __forceinline__ Fast fast(const ProgramDev &pd, const unsigned char *smem) {
    Fast f;
    if constexpr (kByte) {
      f.tBase = smem_u32(smem + kClsBytes);
    }
    else if constexpr (kPtx) {
      f.clsBase = smem_u32(smem);
      f.tBase = smem_u32(smem + kClsBytes);
    }
    else {
      f.tab = table(pd, smem);
      f.cls = reinterpret_cast<const ClsT *>(smem);
    }
    return f;
  }
  // stages the tables; returns the aligned base.  Ends with __syncthreads().
  static __device__ __forceinline__ unsigned char *stage(const ProgramDev &pd,
                                                         unsigned char *smemRaw) {
    unsigned char *smem = smemRaw + ((256u - (smem_u32(smemRaw) & 255u)) & 255u);
    stage_classes<ClsT>(pd, smem);
    if (SMEM)
      stage_table(pd, smem + kClsBytes);
    else if (pd.hotBytes) {
      // mirror of the hottest rows of the global table (a multiple of 16 bytes)
      const uint4 *src = reinterpret_cast<const uint4 *>(pd.table);
      uint4 *d = reinterpret_cast<uint4 *>(smem + kClsBytes);
      for (uint32_t ii = threadIdx.x; ii < (pd.hotBytes >> 4); ii += blockDim.x)
        d[ii] = __ldg(src + ii);
    }
    __syncthreads();
    return smem;
  }
};

} // namespace dev
} // namespace red_b200
