/* hot_walk.cuh - device code shared by the kernels that walk on the hot image
   (program.h: HotImage; kernels_hot.cu, stream_match.cu): PTX helpers, one
   step, the fast passes over a 16-byte chunk, the careful pass, the re-aligning
   text feed.  Everything here is inline device code; the includer pulls it into
   its own anonymous namespace. */

#pragma once

#include "hot.h"
#include "program.h"
#include "walk.cuh"

namespace red_b200 {
namespace {

using namespace dev;

constexpr unsigned kFullMask = 0xffffffffu;

// ---- PTX helpers ----------------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// the same question without waiting
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// as mbar_try_wait, but the hardware may keep the thread suspended for up to `ns`
__device__ __forceinline__ bool mbar_try_wait_for(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk-async copy global -> shared; completion counted on `bar` in bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 16 bytes global -> shared without a register in between (LDGSTS); srcBytes = 0
// fills the destination with zeros and reads nothing
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t srcBytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t a) {
  uint32_t r;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(static_cast<uint16_t>(v)) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// (AnyTable - the full table wherever it lives, the cold path - is in walk.cuh)

/* Walks n bytes at p (generic pointer: a slot or global memory) on the full
   table with the reference's per-byte decisions (walk.cuh).  True = the walk is
   over; wk.mstart then holds the number of bytes consumed.  Rare: lanes that
   left the hot set, lines longer than a slot. */
template <int STYLE>
__device__ __noinline__ bool cold_bytes(const ProgramDev &pd, uint32_t wide, Walk<STYLE, false> &wk,
                                        const uint8_t *p, uint32_t n, uint32_t pos) {
  const AnyTable tab = any_table(pd, wide);
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t co = __ldg(pd.clsOffset + p[i]);
    if (wk.step(pd, tab, co, pos + i)) {
      wk.mstart = pos + i + 1;
      return true;
    }
  }
  return false;
}

/* One byte on the full table (a lane that left the hot set); true = the walk is
   over. */
template <int STYLE>
__device__ __noinline__ bool cold_step(const ProgramDev &pd, uint32_t wide, Walk<STYLE, false> &wk,
                                       uint32_t byte, uint32_t pos) {
  const AnyTable tab = any_table(pd, wide);
  return wk.step(pd, tab, __ldg(pd.clsOffset + byte), pos);
}

struct HotCtx {
  uint32_t tab;     // shared address of the hot table
  uint32_t c2v;     // shared address of codeToV; codeToResult sits 1024 bytes behind it
  uint32_t sink;    // = nHot
  uint32_t cLimit;  // codes >= cLimit need a decision
  uint32_t cAbs, cPure;
  uint32_t wide;
  const uint8_t *codeOf;
};

// hx.tab has a zero low byte, so PRMT puts a text byte and the table base
// together in one instruction: a step is PRMT, IMAD, LDS.U8
template <int STRIDE>
__device__ __forceinline__ uint32_t hot_next(const HotCtx &hx, uint32_t c, uint32_t word, int jj) {
  return lds8(c * STRIDE + __byte_perm(word, hx.tab, 0x7650 | jj));
}
template <int STRIDE>
__device__ __forceinline__ uint32_t hot_next_byte(const HotCtx &hx, uint32_t c, uint32_t byte) {
  return lds8(c * STRIDE + hx.tab + byte);
}

/* A lane's line starts at any byte of its slot, but the lanes of a tile are to
   walk in step *by position in their lines* (byte 0 of every line in the same
   instruction: lines of a log look alike at equal positions, and lanes in like
   states on like bytes share shared-memory words instead of colliding in the
   banks - measured on the union64 corpus: 2.0 -> 1.3 wavefronts per lookup).
   The bulk copy cannot shift a line (source and destination of cp.async.bulk
   are 16-byte aligned, and the tensor form faults on a coordinate that is not:
   tools/micro/tma_lines.cu), so the lane re-aligns in registers: two aligned
   16-byte blocks, a two-level select by the word offset, a byte permute by the
   byte offset - 16 ALU operations per 16 bytes, none in the load/store pipe. */
struct Feed {
  uint32_t blk;   // address of the next aligned block to load
  uint4 cur;      // the two blocks in front of it: the next chunk's bytes are in these
  uint4 nxt;
  uint32_t sel;   // PRMT selector of the byte offset: 0x3210 + 0x1111 * (start & 3)
  bool w1, w2;    // bits of the word offset (start >> 2) & 3
};
__device__ __forceinline__ void feed_begin(Feed &f, const uint32_t textAddr) {
  f.blk = textAddr & ~15u;
  f.cur = lds128(f.blk);
  f.nxt = lds128(f.blk + 16);
  f.blk += 32;
  f.sel = 0x3210u + 0x1111u * (textAddr & 3u);
  f.w1 = (textAddr & 4u) != 0;
  f.w2 = (textAddr & 8u) != 0;
}
// 16 bytes from two aligned blocks, by the word and byte offsets in f
__device__ __forceinline__ uint4 feed_align(const Feed &f, const uint4 &c, const uint4 &n) {
  const uint32_t u0 = f.w1 ? c.y : c.x, u1 = f.w1 ? c.z : c.y, u2 = f.w1 ? c.w : c.z,
                 u3 = f.w1 ? n.x : c.w, u4 = f.w1 ? n.y : n.x, u5 = f.w1 ? n.z : n.y, u6 = f.w1 ? n.w : n.z;
  const uint32_t y0 = f.w2 ? u2 : u0, y1 = f.w2 ? u3 : u1, y2 = f.w2 ? u4 : u2, y3 = f.w2 ? u5 : u3,
                 y4 = f.w2 ? u6 : u4;
  uint4 r;
  r.x = __byte_perm(y0, y1, f.sel);
  r.y = __byte_perm(y1, y2, f.sel);
  r.z = __byte_perm(y2, y3, f.sel);
  r.w = __byte_perm(y3, y4, f.sel);
  return r;
}
/* The next 16 bytes of the line (whatever follows the line's end comes along).
   The block the chunk BEHIND this one will need is asked for first: the loads
   are volatile asm and keep their order, so a load placed in front of a
   chunk's sixteen lookups is under way while they run, and the next chunk's
   first lookup does not wait for its text (it used to: an LDS.128 round trip
   and the re-alignment in front of every chunk's chain).  A feed therefore
   reads up to two blocks past its line's last chunk (launch_batch_hot leaves
   the room behind the last slot). */
__device__ __forceinline__ uint4 feed_next(Feed &f) {
  const uint4 ahead = lds128(f.blk);
  f.blk += 16;
  const uint4 r = feed_align(f, f.cur, f.nxt);
  f.cur = f.nxt;
  f.nxt = ahead;
  return r;
}
/* The fast pass over one aligned 16-byte chunk of one line per lane: no
   decisions, only the largest code per quarter and the code at each quarter's
   end.  [lo, hi): the bytes of the chunk that belong to the line; FULL16: they
   all do.  Returns false when no code reached hx.cLimit (code is then the
   state behind the chunk); true when one did: (j, cj) is where the careful
   pass has to start, code is unchanged. */
template <bool FULL16, int STRIDE>
__device__ __forceinline__ bool hot_fast(const HotCtx &hx, const uint4 &w, const uint32_t lo, const uint32_t hi,
                                         uint32_t &code, uint32_t &j, uint32_t &cj) {
  uint32_t c = code;
  uint32_t mq0 = 0, mq1 = 0, mq2 = 0, mq3 = 0;
  uint32_t cp0 = 0, cp1 = 0, cp2 = 0;
  const uint32_t mask = FULL16 ? 0xffffu : ((0xffffu >> (16 - hi)) & (0xffffu << lo));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t word = (q == 0) ? w.x : (q == 1) ? w.y : (q == 2) ? w.z : w.w;
    uint32_t m = 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      // a byte outside the line: no lookup at all (a predicated load), the
      // lane's lookups would only collide with the others' in the banks
      if (FULL16 || (mask & (1u << (4 * q + jj))))
        c = hot_next<STRIDE>(hx, c, word, jj);
      m = max(m, c);
    }
    if (q == 0) { mq0 = m; cp0 = c; }
    if (q == 1) { mq1 = m; cp1 = c; }
    if (q == 2) { mq2 = m; cp2 = c; }
    if (q == 3) { mq3 = m; }
  }
  // rows of 128 columns: a byte >= 128 anywhere in the chunk and the lookups
  // above mean nothing (they stayed inside shared memory: codes are bytes)
  const bool high = (STRIDE == static_cast<int>(kHotStrideAscii)) && (((w.x | w.y | w.z | w.w) & 0x80808080u) != 0);
  if (!high && (max(max(mq0, mq1), max(mq2, mq3)) < hx.cLimit)) {
    code = c;
    return false;
  }
  // the careful pass starts at the first flagged quarter
  j = 0;
  cj = code;
  if (high) {
    j = lo;
    return true;
  }
  if (mq0 < hx.cLimit) {
    j = 4; cj = cp0;
    if (mq1 < hx.cLimit) {
      j = 8; cj = cp1;
      if (mq2 < hx.cLimit) { j = 12; cj = cp2; }
    }
  }
  j = max(j, lo);
  return true;
}

/* The fast pass where codes that need a decision STICK (style Full: absorbing
   and dead states, the sink; style Instant on an image built with
   acceptSticks): no running maximum - the code behind the chunk tells whether
   anything happened in it.  Returns that code, or 255 for a chunk of a
   128-column image that holds a byte >= 128.  hi: bytes of the chunk that
   belong to the line (FULL16: all 16; the others are not looked up). */
template <bool FULL16, int STRIDE>
__device__ __forceinline__ uint32_t hot_run(const HotCtx &hx, const uint4 &w, const uint32_t hi, uint32_t c) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t word = (q == 0) ? w.x : (q == 1) ? w.y : (q == 2) ? w.z : w.w;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (FULL16 || (static_cast<uint32_t>(4 * q + jj) < hi))
        c = hot_next<STRIDE>(hx, c, word, jj);
  }
  // (rows of 128 columns: whatever follows the line's end in the chunk is looked
  // at too - a false alarm costs a careful pass, not an answer)
  if ((STRIDE == static_cast<int>(kHotStrideAscii)) && (((w.x | w.y | w.z | w.w) & 0x80808080u) != 0))
    return 255u;
  return c;
}

/* hot_run that also hands back the codes behind the chunk's first three
   quarters (no extra instructions, three registers): where deciding codes
   stick they tell in which quarter the run met one, and the second look
   (hot_locate_quarter, or the careful pass) starts there instead of at the
   chunk's first byte. */
template <bool FULL16, int STRIDE>
__device__ __forceinline__ uint32_t hot_run_q(const HotCtx &hx, const uint4 &w, const uint32_t hi, uint32_t c,
                                              uint32_t &cp0, uint32_t &cp1, uint32_t &cp2) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t word = (q == 0) ? w.x : (q == 1) ? w.y : (q == 2) ? w.z : w.w;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (FULL16 || (static_cast<uint32_t>(4 * q + jj) < hi))
        c = hot_next<STRIDE>(hx, c, word, jj);
    if (q == 0) cp0 = c;
    if (q == 1) cp1 = c;
    if (q == 2) cp2 = c;
  }
  if ((STRIDE == static_cast<int>(kHotStrideAscii)) && (((w.x | w.y | w.z | w.w) & 0x80808080u) != 0))
    return 255u;
  return c;
}

// hot_locate over one quarter of a chunk: word = its four bytes, hi of them in the line
template <int STRIDE>
__device__ __forceinline__ uint32_t hot_locate_quarter(const HotCtx &hx, const uint32_t word, const uint32_t hi, uint32_t c) {
  uint32_t n = 0;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj)
    if (static_cast<uint32_t>(jj) < hi) {
      c = hot_next<STRIDE>(hx, c, word, jj);
      n += (c < hx.cLimit) ? 1u : 0u;
    }
  return n;
}

/* Behind hot_run on an image whose accepting codes stick: how many bytes of
   the chunk come before the one that entered the code the run ended in (the
   same 16 lookups again, counting the codes below the limit). */
template <int STRIDE>
__device__ __forceinline__ uint32_t hot_locate(const HotCtx &hx, const uint4 &w, const uint32_t hi, uint32_t c) {
  uint32_t n = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t word = (q == 0) ? w.x : (q == 1) ? w.y : (q == 2) ? w.z : w.w;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (static_cast<uint32_t>(4 * q + jj) < hi) {
        c = hot_next<STRIDE>(hx, c, word, jj);
        n += (c < hx.cLimit) ? 1u : 0u;
      }
  }
  return n;
}

/* The careful pass: bytes [j, hi) of the chunk at chunkAddr one by one with
   the style's decisions; on the hot table while the state is hot, on the full
   table while it is not (First/Tangent are not walked here: the limit never
   changes).  c: the state in front of byte j in, behind the last byte taken
   out (hx.sink: it is in wk.v).  Returns true when the walk is over. */
template <int STYLE, int STRIDE, bool ACC_AS_V = false>
__device__ __forceinline__ bool hot_careful(const ProgramDev &pd, const HotCtx &hx, const uint32_t chunkAddr,
                                            uint32_t j, const uint32_t hi, const uint32_t pos0,
                                            uint32_t &c, Walk<STYLE, false> &wk) {
#pragma unroll 1
  for (; j < hi; ++j) {
    const uint32_t byte = lds8(chunkAddr + j);
    bool stopped = false;
    if (c == hx.sink) {
      // (a copy goes through the call, so that wk itself can stay in registers)
      Walk<STYLE, false> tmp = wk;
      stopped = cold_step<STYLE>(pd, hx.wide, tmp, byte, pos0 + j);
      wk = tmp;
      if (!stopped)
        c = hx.codeOf[state_index(pd, wk.v)];
    }
    else {
      const uint32_t nc = ((STRIDE == static_cast<int>(kHotStrideAscii)) && (byte >= 128))
          ? hx.sink : hot_next_byte<STRIDE>(hx, c, byte);
      if (nc >= hx.cLimit) {
        if (nc == hx.sink) {
          // leaves the hot set: this byte again, on the full table
          wk.v = lds32(hx.c2v + 4 * c);
          c = hx.sink;
          --j;
          continue;
        }
        if (nc >= hx.cPure)
          stopped = true;  // pure dead end (Matcher.h:478)
        else if (STYLE == STY_FULL)
          stopped = true;  // absorbing accept: the outcome is known
        else {
          // the last accepting state: tagged, the result id itself - or (the
          // stream matcher stitches chunks by states) its v
          wk.accV = ACC_AS_V ? lds32(hx.c2v + 4 * nc) : (0x80000000u | lds32(hx.c2v + 1024 + 4 * nc));
          wk.mend = pos0 + j + 1;
          if (STYLE == STY_INSTANT)
            stopped = true;
          else if (nc >= hx.cAbs) {
            wk.absorbed = true;
            stopped = true;
          }
        }
      }
      c = nc;
    }
    if (stopped) {
      wk.mstart = pos0 + j + 1;  // bytes consumed
      return true;
    }
  }
  return false;
}

// fast pass, and the careful pass right behind it where it is needed
template <int STYLE, bool FULL16, int STRIDE, bool ACC_AS_V = false>
__device__ __forceinline__ bool hot_chunk(const ProgramDev &pd, const HotCtx &hx, const uint4 &w,
                                          const uint32_t chunkAddr, const uint32_t lo, const uint32_t hi,
                                          const uint32_t pos0, uint32_t &code, Walk<STYLE, false> &wk) {
  uint32_t j = lo, cj = code;
  if ((code != hx.sink) && !hot_fast<FULL16, STRIDE>(hx, w, lo, hi, code, j, cj))
    return false;
  const bool stopped = hot_careful<STYLE, STRIDE, ACC_AS_V>(pd, hx, chunkAddr, j, hi, pos0, cj, wk);
  code = cj;
  return stopped;
}


} // anonymous
} // namespace red_b200
