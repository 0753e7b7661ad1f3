/* program.cpp - parse/validate a serialized RED program, build the GPU image.
   See program.h for both layouts.  Host code only (no CUDA). */

#include "program.h"

#include <algorithm>
#include <cstring>

namespace red_b200 {

namespace {

template <class T> T loadLE(const uint8_t *p) {
  T v;
  memcpy(&v, p, sizeof(T));
  return v;
}

uint64_t readValue(const uint8_t *p, uint32_t width) {
  switch (width) {
  case 1: return *p;
  case 2: return loadLE<uint16_t>(p);
  default: return loadLE<uint32_t>(p);
  }
}

} // anonymous

// FNV-1a (32 bit) of everything from the format byte on
// (reference: calcChecksum, lib/Serializer.cpp:301; include/Fnv.h:45-63)
uint32_t calc_checksum(const void *ptr, size_t len) {
  const uint8_t *p = static_cast<const uint8_t *>(ptr);
  uint32_t h = 0x811c9dc5u;
  for (size_t ii = 12; ii < len; ++ii)
    h = (h ^ p[ii]) * 0x01000193u;
  return h;
}

// reference: checkHeader, lib/Serializer.cpp:270-298 (same order, same text)
const char *check_header(const void *ptr, size_t len) {
  if (!ptr || (len < kHeaderSize))
    return "Serialized DFA: header too short";
  const uint8_t *p = static_cast<const uint8_t *>(ptr);
  if (memcmp(p, "REDA", 4) != 0)
    return "Serialized DFA: bad magic number";
  if ((loadLE<uint16_t>(p + 4) != 1) || (loadLE<uint16_t>(p + 6) != 0))
    return "Serialized DFA: unrecognized version";
  uint32_t stored = loadLE<uint32_t>(p + 8);
  uint32_t actual = calc_checksum(ptr, len);
  if (stored != actual) {
    if (stored == __builtin_bswap32(actual))
      return "serialized DFA: foreign endian-ness";
    return "serialized DFA: checksum mismatch";
  }
  uint8_t fmt = p[12];
  if ((fmt != 1) && (fmt != 2) && (fmt != 4))
    return "Serialized DFA: unsupported format";
  return nullptr;
}

bool parse_program(const void *ptr, size_t len, ParsedProgram &out,
                   std::string &why) {
  if (const char *msg = check_header(ptr, len)) {
    why = msg;
    return false;
  }
  const uint8_t *p = static_cast<const uint8_t *>(ptr);
  out.format = p[12];
  out.maxChar = p[13];
  out.classCount = out.maxChar + 1;
  out.leaderLen = p[14];
  out.stateCount = loadLE<uint32_t>(p + 16);
  uint32_t initialOff = loadLE<uint32_t>(p + 20);
  uint32_t leaderOff = loadLE<uint32_t>(p + 24);
  memcpy(out.equiv, p + 32, 256);

  // The reference trusts the checksum; the GPU must never index out of
  // bounds, so the structure is verified as well.
  const size_t width = out.format;
  const size_t rowVals = static_cast<size_t>(out.maxChar) + 2;
  const size_t pad = (out.leaderLen + 7u) & ~7u;
  const size_t need = kHeaderSize + pad +
                      static_cast<size_t>(out.stateCount) * rowVals * width;
  if (out.stateCount < 2) {
    why = "serialized DFA: fewer than two states";
    return false;
  }
  if (len < need) {
    why = "serialized DFA: truncated state table";
    return false;
  }
  for (int ii = 0; ii < 256; ++ii)
    if (out.equiv[ii] > out.maxChar) {
      why = "serialized DFA: class map exceeds maxChar";
      return false;
    }
  out.leader.assign(p + kHeaderSize, p + kHeaderSize + out.leaderLen);
  for (uint8_t c : out.leader)
    if (c > out.maxChar) {
      why = "serialized DFA: leader exceeds maxChar";
      return false;
    }

  auto offToRow = [&](uint64_t off, uint32_t &row) {
    if ((off % rowVals) != 0)
      return false;
    uint64_t r = off / rowVals;
    if (r >= out.stateCount)
      return false;
    row = static_cast<uint32_t>(r);
    return true;
  };
  // initialOff_/leaderOff_ are BYTE offsets from the first row (they are fed
  // to DfaProxy::init, Proxy.h:127), unlike the row entries, which are in
  // value units (DfaProxy::trans multiplies by sizeof(Value), Proxy.h:143).
  if ((initialOff % width) || (leaderOff % width) ||
      !offToRow(initialOff / width, out.initialState) ||
      !offToRow(leaderOff / width, out.leaderState)) {
    why = "serialized DFA: initial/leader offset out of range";
    return false;
  }

  const uint8_t *base = p + kHeaderSize + pad;
  const uint32_t topBit = 8 * out.format - 1;
  const uint64_t deadMask = 1ull << topBit;
  const uint32_t C = out.classCount;
  out.result.resize(out.stateCount);
  out.deadEnd.resize(out.stateCount);
  out.pureDead.resize(out.stateCount);
  out.next.resize(static_cast<size_t>(out.stateCount) * C);
  for (uint32_t st = 0; st < out.stateCount; ++st) {
    const uint8_t *row = base + static_cast<size_t>(st) * rowVals * width;
    uint64_t head = readValue(row, out.format);
    out.result[st] = static_cast<int32_t>(head & (deadMask - 1));
    out.deadEnd[st] = (head >> topBit) & 1;
    out.pureDead[st] = (head == deadMask);
    for (uint32_t c = 0; c < C; ++c) {
      uint64_t off = readValue(row + (1 + c) * width, out.format);
      uint32_t nx;
      if (!offToRow(off, nx)) {
        why = "serialized DFA: transition out of range";
        return false;
      }
      out.next[static_cast<size_t>(st) * C + c] = nx;
    }
  }
  return true;
}

void build_image(const ParsedProgram &pp, size_t smemBudget,
                 DeviceImage &img, bool allowByteRows, const uint64_t *visitsByOld) {
  const uint32_t n = pp.stateCount;
  const uint32_t C = pp.classCount;

  // renumber: rejecting states, then accepting, then pure dead ends, so that
  // the per-byte test for "nothing special happened" is one compare:
  //   v < acceptV            ordinary rejecting state (the common case)
  //   acceptV <= v < pureV   accepting (absV <= v: and can never be left)
  //   v >= pureV             pure dead end (reference: DfaProxy::pureDeadEnd)
  std::vector<uint32_t> order;
  order.reserve(n);
  for (uint32_t st = 0; st < n; ++st)
    if (!pp.pureDead[st] && (pp.result[st] <= 0))
      order.push_back(st);
  // An accepting state all of whose transitions lead back to itself (the tail
  // of a pattern with a loose end, ".*foo.*") can never be left: once there,
  // the outcome of every style is known without reading the rest of the line.
  auto selfLoop = [&](uint32_t st) {
    for (uint32_t c = 0; c < C; ++c)
      if (pp.next[static_cast<size_t>(st) * C + c] != st)
        return false;
    return true;
  };
  uint32_t firstAccept = static_cast<uint32_t>(order.size());
  for (uint32_t st = 0; st < n; ++st)
    if (!pp.pureDead[st] && (pp.result[st] > 0) && !selfLoop(st))
      order.push_back(st);
  uint32_t firstAbsorb = static_cast<uint32_t>(order.size());
  for (uint32_t st = 0; st < n; ++st)
    if (!pp.pureDead[st] && (pp.result[st] > 0) && selfLoop(st))
      order.push_back(st);
  uint32_t firstPure = static_cast<uint32_t>(order.size());
  for (uint32_t st = 0; st < n; ++st)
    if (pp.pureDead[st])
      order.push_back(st);
  if (visitsByOld) {
    auto byVisits = [&](uint32_t a, uint32_t b) { return visitsByOld[a] > visitsByOld[b]; };
    std::stable_sort(order.begin(), order.begin() + firstAccept, byVisits);
    std::stable_sort(order.begin() + firstAccept, order.begin() + firstAbsorb, byVisits);
    std::stable_sort(order.begin() + firstAbsorb, order.begin() + firstPure, byVisits);
  }
  img.acceptCount = firstPure - firstAccept;
  img.pureCount = n - firstPure;
  std::vector<uint32_t> oldToNew(n);
  for (uint32_t ii = 0; ii < n; ++ii)
    oldToNew[order[ii]] = ii;
  img.newToOld = order;
  img.oldToNew = oldToNew;

  img.stateCount = n;
  img.classCount = C;
  img.results.resize(n);
  img.maxResult = 0;
  for (uint32_t ii = 0; ii < n; ++ii) {
    img.results[ii] = pp.result[order[ii]];
    img.maxResult = std::max(img.maxResult, img.results[ii]);
  }
  img.initResult = pp.result[pp.initialState];
  img.leader = pp.leader;

  // can a walk from the initial state ever reach a state it cannot leave
  // (absorbing accept or pure dead end)?  The long-stream matcher uses this:
  // if not, nothing can decide a Last/Full walk before the end of the text.
  {
    std::vector<char> seen(n, 0);
    std::vector<uint32_t> todo{pp.initialState};
    seen[pp.initialState] = 1;
    img.finalReachable = false;
    while (!todo.empty()) {
      const uint32_t st = todo.back();
      todo.pop_back();
      if (oldToNew[st] >= firstAbsorb) {
        img.finalReachable = true;
        break;
      }
      for (uint32_t c = 0; c < C; ++c) {
        const uint32_t nx = pp.next[static_cast<size_t>(st) * C + c];
        if (!seen[nx]) {
          seen[nx] = 1;
          todo.push_back(nx);
        }
      }
    }
  }

  // The reference stops walking the moment it enters a pure dead end
  // (Matcher.h:405, 478) and never looks at that row's exits.  Making those
  // rows absorbing in the image is the same semantics, and lets a kernel step
  // through them without a per-byte check.
  auto absorbing = [&](uint32_t newIdx) { return newIdx >= firstPure; };

  // Small DFAs: rows of 256 entries indexed by the raw byte, so that a step
  // is ONE shared-memory load (no class lookup).  v is the state index.
  const bool byteRows = allowByteRows &&
                        (static_cast<uint64_t>(n) * 256 * 2 <= smemBudget) &&
                        (n <= 0xffff);

  // otherwise: smallest unit shift for which every v fits 16 bits
  bool smem = false;
  uint32_t shift = 0, vps = 0, stride = 0;
  for (shift = 0; !byteRows && (shift <= 1); ++shift) {
    vps = (C + (1u << shift) - 1) >> shift;
    stride = vps << shift;
    uint64_t maxV = static_cast<uint64_t>(n) * vps;  // acceptV may equal n*vps
    uint64_t bytes = static_cast<uint64_t>(n) * stride * 2;
    if ((maxV <= 0xffff) && (bytes <= smemBudget)) {
      smem = true;
      break;
    }
    if (bytes > smemBudget)
      break;
  }

  img.clsOffset.resize(256);
  img.byteRows = byteRows;
  if (byteRows) {
    img.mode = kTableSmemU16;
    img.unitShift = 8;
    img.vPerState = 1;
    img.rowStride = 256;
    img.table16.assign(static_cast<size_t>(n) * 256, 0);
    for (uint32_t ii = 0; ii < n; ++ii) {
      const uint32_t *src = &pp.next[static_cast<size_t>(order[ii]) * C];
      uint16_t *dst = &img.table16[static_cast<size_t>(ii) * 256];
      for (uint32_t b = 0; b < 256; ++b)
        dst[b] = static_cast<uint16_t>(absorbing(ii) ? ii : oldToNew[src[pp.equiv[b]]]);
    }
    for (int b = 0; b < 256; ++b)
      img.clsOffset[b] = static_cast<uint16_t>(b * 2);  // identity "classes"
  }
  else if (smem) {
    img.mode = kTableSmemU16;
    img.unitShift = shift;
    img.vPerState = vps;
    img.rowStride = stride;
    img.table16.assign(static_cast<size_t>(n) * stride, 0);
    for (uint32_t ii = 0; ii < n; ++ii) {
      const uint32_t *src = &pp.next[static_cast<size_t>(order[ii]) * C];
      uint16_t *dst = &img.table16[static_cast<size_t>(ii) * stride];
      for (uint32_t c = 0; c < C; ++c)
        dst[c] = static_cast<uint16_t>((absorbing(ii) ? ii : oldToNew[src[c]]) * vps);
      // padding entries are never addressed (class < C); point them at self
      for (uint32_t c = C; c < stride; ++c)
        dst[c] = static_cast<uint16_t>(ii * vps);
    }
    for (int b = 0; b < 256; ++b)
      img.clsOffset[b] = static_cast<uint16_t>(pp.equiv[b] * 2);
  }
  else {
    img.mode = kTableGlobalU32;
    img.unitShift = 0;
    img.vPerState = C;
    img.rowStride = C;
    img.table32.resize(static_cast<size_t>(n) * C);
    for (uint32_t ii = 0; ii < n; ++ii) {
      const uint32_t *src = &pp.next[static_cast<size_t>(order[ii]) * C];
      uint32_t *dst = &img.table32[static_cast<size_t>(ii) * C];
      for (uint32_t c = 0; c < C; ++c)
        dst[c] = (absorbing(ii) ? ii : oldToNew[src[c]]) * C;
    }
    for (int b = 0; b < 256; ++b)
      img.clsOffset[b] = static_cast<uint16_t>(pp.equiv[b] * 4);
  }
  // images, under one byte, of the set of states a walk from the initial
  // state can be in at all (long-stream matcher)
  {
    img.resetCount.assign(256, 0);
    img.resetStates.assign(256 * DeviceImage::kResetMax, 0);
    std::vector<char> reach(n, 0);  // by new index
    {
      std::vector<uint32_t> todo{oldToNew[pp.initialState]};
      reach[todo[0]] = 1;
      while (!todo.empty()) {
        const uint32_t ii = todo.back();
        todo.pop_back();
        if (absorbing(ii))
          continue;
        for (uint32_t c = 0; c < C; ++c) {
          const uint32_t to = oldToNew[pp.next[static_cast<size_t>(order[ii]) * C + c]];
          if (!reach[to]) {
            reach[to] = 1;
            todo.push_back(to);
          }
        }
      }
    }
    std::vector<uint32_t> stamp(n, 0xffffffffu);
    for (uint32_t c = 0; c < C; ++c) {
      uint32_t found[DeviceImage::kResetMax];
      uint32_t cnt = 0;
      bool small = true;
      for (uint32_t ii = 0; (ii < n) && small; ++ii) {
        if (!reach[ii])
          continue;
        // (pure dead ends stay where they are, as in the table)
        const uint32_t to = absorbing(ii) ? ii : oldToNew[pp.next[static_cast<size_t>(order[ii]) * C + c]];
        if (stamp[to] != c) {
          stamp[to] = c;
          if (cnt == DeviceImage::kResetMax)
            small = false;
          else
            found[cnt++] = to;
        }
      }
      if (!small)
        continue;
      std::sort(found, found + cnt);
      for (int b = 0; b < 256; ++b)
        if (pp.equiv[b] == c) {
          img.resetCount[b] = static_cast<uint8_t>(cnt);
          for (uint32_t k = 0; k < cnt; ++k)
            img.resetStates[b * DeviceImage::kResetMax + k] = found[k] * img.vPerState;
        }
    }
  }
  img.initV = oldToNew[pp.initialState] * img.vPerState;
  img.leaderV = oldToNew[pp.leaderState] * img.vPerState;
  img.pureV = firstPure * img.vPerState;
  img.acceptV = firstAccept * img.vPerState;
  img.absV = firstAbsorb * img.vPerState;
}

void build_hot_image(const ParsedProgram &pp, const DeviceImage &img,
                     const uint64_t *visits, uint32_t maxHot, bool asciiRows, bool acceptSticks,
                     HotImage &hot, bool wideCodes) {
  const uint32_t n = pp.stateCount;
  const uint32_t C = pp.classCount;
  maxHot = std::max(1u, std::min(maxHot, wideCodes ? kHotMaxWide : kHotMax));
  const uint32_t firstPure = img.pureV / img.vPerState;
  // successor in renumbered indices; pure dead ends loop to themselves, as in
  // the full image
  auto nextOf = [&](uint32_t idx, uint32_t byte) -> uint32_t {
    if (idx >= firstPure)
      return idx;
    return img.oldToNew[pp.next[static_cast<size_t>(img.newToOld[idx]) * C + pp.equiv[byte]]];
  };
  // breadth-first order from the initial state
  const uint32_t init = img.initV / img.vPerState;
  std::vector<uint32_t> bfs;
  std::vector<uint32_t> rank(n, 0xffffffffu);
  bfs.reserve(n);
  bfs.push_back(init);
  rank[init] = 0;
  for (size_t head = 0; head < bfs.size(); ++head) {
    const uint32_t st = bfs[head];
    if (st >= firstPure)
      continue;
    const uint32_t *row = &pp.next[static_cast<size_t>(img.newToOld[st]) * C];
    for (uint32_t c = 0; c < C; ++c) {
      const uint32_t nx = img.oldToNew[row[c]];
      if (rank[nx] == 0xffffffffu) {
        rank[nx] = static_cast<uint32_t>(bfs.size());
        bfs.push_back(nx);
      }
    }
  }
  // the hottest states: by visits (ties and the unvisited by depth)
  std::vector<uint32_t> cand = bfs;
  if (visits)
    std::stable_sort(cand.begin(), cand.end(), [&](uint32_t a, uint32_t b) {
      if ((a == init) != (b == init))
        return a == init;
      return visits[a] > visits[b];
    });
  if (visits) {
    // what the sample never entered stays on the full table
    size_t seen = 1;
    while ((seen < cand.size()) && (visits[cand[seen]] > 0))
      ++seen;
    cand.resize(seen);
  }
  if (cand.size() > maxHot)
    cand.resize(maxHot);
  std::sort(cand.begin(), cand.end());  // codes follow the state order

  if (visits) {
    uint64_t all = 0, kept = 0;
    for (uint32_t st = 0; st < n; ++st)
      all += visits[st];
    for (uint32_t st : cand)
      kept += visits[st];
    hot.coverage = all ? static_cast<uint32_t>(kept * 1000 / all) : 1000;
  }
  else
    hot.coverage = (cand.size() == bfs.size()) ? 1000 : 0;  // no profile: unknown unless every state is hot
  hot.nHot = static_cast<uint32_t>(cand.size());
  hot.entryBytes = wideCodes ? 2 : 1;
  std::vector<uint32_t> codeOf(n, hot.nHot);
  hot.codeToV.assign(std::max<size_t>(256, hot.nHot + 1), 0);
  for (uint32_t c = 0; c < hot.nHot; ++c) {
    codeOf[cand[c]] = c;
    hot.codeToV[c] = cand[c] * img.vPerState;
  }
  auto below = [&](uint32_t v) {
    const uint32_t idx = v / img.vPerState;
    return static_cast<uint32_t>(std::lower_bound(cand.begin(), cand.end(), idx) - cand.begin());
  };
  hot.cAccept = below(img.acceptV);
  hot.cAbs = below(img.absV);
  hot.cPure = below(img.pureV);
  hot.initCode = codeOf[init];
  // (a walk that STARTS in an accepting state leaves it by its real row)
  if (hot.initCode >= hot.cAccept)
    acceptSticks = false;
  const uint32_t columns = asciiRows ? 128 : 256;
  const uint32_t E = hot.entryBytes;
  hot.stride = columns * E + 4;
  const size_t bytes = (static_cast<size_t>(hot.nHot + 1) * hot.stride + 15) & ~size_t(15);
  hot.table.assign(bytes, 0);
  auto put = [&](uint8_t *at, uint32_t code) {
    at[0] = static_cast<uint8_t>(code);
    if (E == 2)
      at[1] = static_cast<uint8_t>(code >> 8);
  };
  for (uint32_t c = 0; c <= hot.nHot; ++c) {
    uint8_t *row = &hot.table[static_cast<size_t>(c) * hot.stride];
    for (uint32_t b = 0; b < columns; ++b)
      put(row + b * E, ((c == hot.nHot) || (acceptSticks && (c >= hot.cAccept))) ? c : codeOf[nextOf(cand[c], b)]);
    // the bytes that pad a row to an odd number of banks hold its own code
    for (uint32_t b = columns * E; b < hot.stride; b += E)
      put(row + b, c);
  }
  for (size_t i = static_cast<size_t>(hot.nHot + 1) * hot.stride; i + E <= bytes; i += E)
    put(&hot.table[i], hot.nHot);
  hot.codeOf.assign(static_cast<size_t>(n) * E, 0);
  for (uint32_t st = 0; st < n; ++st)
    put(&hot.codeOf[static_cast<size_t>(st) * E], codeOf[st]);
  hot.acceptSticks = acceptSticks;
}

} // namespace red_b200
