/* walk.cuh - device code shared by the batch kernels: table accessors, the
   per-byte walk with the reference's style semantics, table staging. */

#pragma once

#include "device_program.h"

namespace red_b200 {
namespace dev {

constexpr uint32_t kNone = 0xffffffffu;
constexpr int kMaxThreads = 1024;

enum { STY_INSTANT = 1, STY_FIRST = 2, STY_TANGENT = 3, STY_LAST = 4,
       STY_FULL = 5 };

// ---- table accessors ------------------------------------------------------

// u16 entries in shared memory; K = unitShift + 1 (byte offset of row = v<<K)
template <int K> struct SmemTable {
  const unsigned char *base;
  __device__ __forceinline__ uint32_t next(uint32_t v, uint32_t co) const {
    return *reinterpret_cast<const uint16_t *>(base + ((v << K) + co));
  }
};

// u32 entries in global memory (L2 resident); byte offset of row = v * 4.
// The first hotBytes of the table - the rows of the states nearest the
// initial state, where a walk over real text spends nearly all of its time -
// may be mirrored in shared memory (hotBytes = 0: no mirror).  A lookup in L1
// costs a wavefront per distinct 128-byte line among the lanes, in shared
// memory only per bank conflict.
struct GlobalTable {
  const unsigned char *base;
  uint32_t hotAddr = 0;   // shared-memory address of the mirror
  uint32_t hotBytes = 0;
  __device__ __forceinline__ uint32_t next(uint32_t v, uint32_t co) const {
    const uint32_t off = (v << 2) + co;
    if (off < hotBytes) {
      uint32_t r;
      asm("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(hotAddr + off));
      return r;
    }
    return __ldg(reinterpret_cast<const uint32_t *>(base + off));
  }
};

// the full table in global memory, whichever of the two layouts it has (rare
// paths: lanes that left the hot set, repairs of the long-stream matcher)
struct AnyTable {
  const unsigned char *base;
  uint32_t shift;  // byte offset of a row = v << shift
  uint32_t wide;   // 1: u32 entries, 0: u16
  __device__ __forceinline__ uint32_t next(uint32_t v, uint32_t co) const {
    const unsigned char *p = base + ((static_cast<size_t>(v) << shift) + co);
    return wide ? __ldg(reinterpret_cast<const uint32_t *>(p))
                : static_cast<uint32_t>(__ldg(reinterpret_cast<const uint16_t *>(p)));
  }
};

__device__ __forceinline__ AnyTable any_table(const ProgramDev &pd, uint32_t wide) {
  AnyTable t;
  t.base = static_cast<const unsigned char *>(pd.table);
  t.shift = wide ? 2 : pd.unitShift + 1;
  t.wide = wide;
  return t;
}

// renumbered state index of v = v / vPerState; by multiplication when the
// uploader found a reciprocal that is exact for every v of this program
__device__ __forceinline__ uint32_t state_index(const ProgramDev &pd, uint32_t v) {
  return pd.divMagic ? __umulhi(v, pd.divMagic) : v / pd.vPerState;
}

// ---- the walk -------------------------------------------------------------

template <int STYLE, bool WANT_START> struct Walk {
  uint32_t v;       // current state
  uint32_t limit;   // states >= limit need a decision
  uint32_t accV;    // last accepting state seen (kNone: none)
  uint32_t mstart;  // where the initial state was last left
  uint32_t mend;    // one past the last accepting position
  bool absorbed;    // stopped in an absorbing accepting state: the match runs to the end

  __device__ __forceinline__ void begin(const ProgramDev &pd) {
    v = pd.initV;
    // a Full walk only has to notice states that can never be left (to stop
    // early): absorbing accepting states and pure dead ends
    limit = (STYLE == STY_FULL) ? pd.absV : pd.acceptV;
    absorbed = false;
    accV = kNone;
    mstart = 0;
    mend = 0;
  }

  // out-of-line part of a step; true = the walk is over
  __device__ __forceinline__ bool special(const ProgramDev &pd, uint32_t pos) {
    if (v >= pd.pureV)
      return true;  // pure dead end: Matcher.h:405/478
    if (STYLE == STY_FULL)
      return true;  // v >= absV: the final state is already known
    if (v >= pd.acceptV) {
      if (STYLE == STY_FIRST) {
        // stop when the result changes (Matcher.h:455-460)
        if ((accV != kNone) && (v != accV) &&
            (pd.results[state_index(pd, v)] != pd.results[state_index(pd, accV)]))
          return true;
      }
      accV = v;
      mend = pos + 1;
      if (STYLE == STY_INSTANT)
        return true;
      if (v >= pd.absV) {
        absorbed = true;  // every later position accepts with this result
        return true;
      }
      if ((STYLE == STY_FIRST) || (STYLE == STY_TANGENT))
        limit = 0;  // from now on the first rejecting state ends the walk
      return false;
    }
    // ordinary rejecting state after an accept (First/Tangent only)
    return true;
  }

  template <class Tab>
  __device__ __forceinline__ bool step(const ProgramDev &pd, const Tab &tab,
                                       uint32_t co, uint32_t pos) {
    uint32_t nv = tab.next(v, co);
    if (WANT_START) {
      if ((v == pd.initV) && (nv != pd.initV))
        mstart = pos;
    }
    v = nv;
    if (v >= limit)
      return special(pd, pos);
    return false;
  }

  __device__ __forceinline__ void finish(const ProgramDev &pd, uint32_t len,
                                         int32_t &result, uint32_t &start,
                                         uint32_t &end) const {
    int32_t cur = ((v >= pd.acceptV) && (v < pd.pureV))
                      ? pd.results[state_index(pd, v)] : 0;
    int32_t prev = 0;
    if ((STYLE != STY_FULL) && (STYLE != STY_INSTANT) && (accV != kNone))
      prev = pd.results[state_index(pd, accV)];
    result = (prev > 0) ? prev : cur;
    uint32_t e = ((STYLE == STY_FULL) || absorbed) ? len : mend;
    start = (result > 0) ? mstart : 0;
    end = (result > 0) ? e : 0;
  }
};

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int j) {
  return __byte_perm(w, 0, 0x4440 | j);
}

__device__ __forceinline__ uint32_t word_of(const uint4 &q, int i) {
  return (i == 0) ? q.x : (i == 1) ? q.y : (i == 2) ? q.z : q.w;
}

// ---- kernels --------------------------------------------------------------

/* Shared memory layout (dynamic):
     [0, 512)            class table: u8 (pre-scaled offsets < 256) or u16
     [512, 512 + bytes)  transition table (smem mode only)                 */
constexpr uint32_t kClsBytes = 512;

template <class ClsT>
__device__ __forceinline__ void stage_classes(const ProgramDev &pd,
                                              unsigned char *smem) {
  ClsT *cls = reinterpret_cast<ClsT *>(smem);
  for (int ii = threadIdx.x; ii < 256; ii += blockDim.x)
    cls[ii] = static_cast<ClsT>(pd.clsOffset[ii]);
}

__device__ __forceinline__ void stage_table(const ProgramDev &pd,
                                            unsigned char *dst) {
  // tableBytes is padded to a multiple of 16 by the uploader
  const uint4 *src = reinterpret_cast<const uint4 *>(pd.table);
  uint4 *d = reinterpret_cast<uint4 *>(dst);
  const uint32_t n = pd.tableBytes >> 4;
  for (uint32_t ii = threadIdx.x; ii < n; ii += blockDim.x)
    d[ii] = __ldg(src + ii);
}

} // namespace dev
} // namespace red_b200
