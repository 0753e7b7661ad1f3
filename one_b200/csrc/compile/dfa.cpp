/* dfa.cpp - NFA -> DFA -> minimal DFA -> serialized REDA program.

   Stages (reference files for the behaviour each stage has to reproduce):
     determinize   subset construction with end marks; a state's result is the
                   smallest result id whose end mark is enabled there
                   (lib/Powerset.cpp:91-140, Dfa.cpp chopEndMarks:121)
     byte classes  bytes that behave alike in every state of the *unminimized*
                   DFA share a class (lib/Dfa.cpp:236, called from
                   Minimizer.cpp:124 before minimization - the partition is
                   part of the file format and of the leader, so it is taken
                   at the same point)
     minimize      partition refinement from the partition by result
                   (lib/Minimizer.cpp); states 0 (error) and 1 (initial) keep
                   their numbers
     flags, leader dead-end flag and required prefix (lib/Dfa.cpp:30-44, 166)
     serialize     lib/Serializer.cpp:150-240

   The minimal DFA is unique up to renumbering, so the serialized program
   differs from the reference's only in the order of states and classes.
 */

#include "dfa.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <deque>
#include <map>
#include <set>
#include <unordered_map>

#include "../../../include/red_b200.h"
#include "../program.h"

namespace red_b200 {
namespace rx {

namespace {

struct VecHash {
  size_t operator()(const std::vector<NodeId> &v) const {
    uint64_t h = 0xcbf29ce484222325ull;
    for (NodeId x : v) {
      h ^= static_cast<uint32_t>(x);
      h *= 0x100000001b3ull;
    }
    return static_cast<size_t>(h);
  }
};

} // anonymous

/* Subset construction (reference: lib/Powerset.cpp:40-180 builds the same
   automaton with per-byte multi-char sets; the numbering below - error 0,
   initial 1, then in order of discovery by ascending byte - is the same).

   Two things keep it fast on rule sets with thousands of loose-start rules,
   where every subset contains the same few thousand looping nodes:

   * bytes that no edge of the NFA tells apart are stepped once (NFA-level
     byte classes, usually a few dozen);
   * a subset is held as (core, rest): `core` is an interned set of
     PERSISTENT nodes - nodes that loop to themselves on every byte and
     therefore never leave a subset once in - and `rest` the other members.
     What a core contributes to the step on each class is computed once per
     core; per subset only `rest`, a handful of nodes, is walked.  */
RawDfa determinize(const Nfa &nfa, uint64_t budget) {
  RawDfa dfa;
  const size_t nNodes = nfa.size();

  // --- NFA-level byte classes, numbered by first byte
  unsigned clsOf[256];
  std::vector<unsigned> rep;
  {
    std::map<std::vector<uint64_t>, int> distinct;
    for (size_t id = 0; id < nNodes; ++id)
      for (const Edge &e : nfa[static_cast<NodeId>(id)].out)
        if (!e.on.empty())
          distinct.emplace(std::vector<uint64_t>(e.on.w, e.on.w + 4), 0);
    const size_t words = (distinct.size() + 63) / 64;
    std::vector<std::vector<uint64_t>> sig(256, std::vector<uint64_t>(words, 0));
    size_t k = 0;
    for (auto &kv : distinct) {
      for (unsigned c = 0; c < 256; ++c)
        if ((kv.first[c >> 6] >> (c & 63)) & 1)
          sig[c][k >> 6] |= 1ull << (k & 63);
      ++k;
    }
    std::map<std::vector<uint64_t>, unsigned> seen;
    for (unsigned c = 0; c < 256; ++c) {
      auto it = seen.find(sig[c]);
      if (it == seen.end()) {
        it = seen.emplace(sig[c], static_cast<unsigned>(rep.size())).first;
        rep.push_back(c);
      }
      clsOf[c] = it->second;
    }
  }
  const unsigned C = static_cast<unsigned>(rep.size());

  // --- per node: targets per class (CSR), smallest end mark, persistence
  struct Fanout {
    int32_t mark = 0;
    bool persistent = false;
    std::vector<uint32_t> begin;  // C + 1
    std::vector<NodeId> to;
  };
  std::vector<Fanout> fan(nNodes);
  std::vector<char> built(nNodes, 0);
  auto fanout = [&](NodeId id) -> const Fanout & {
    Fanout &f = fan[id];
    if (!built[id]) {
      built[id] = 1;
      std::vector<std::pair<uint32_t, NodeId>> pairs;
      for (const Edge &e : nfa[id].out) {
        if (e.endMark > 0)
          f.mark = f.mark ? std::min(f.mark, e.endMark) : e.endMark;
        if (e.on.empty())
          continue;
        for (unsigned c = 0; c < C; ++c)
          if (e.on.has(rep[c]))
            pairs.emplace_back(c, e.to);
      }
      std::sort(pairs.begin(), pairs.end());
      pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
      f.begin.assign(C + 1, 0);
      for (auto &pr : pairs)
        ++f.begin[pr.first + 1];
      for (unsigned c = 0; c < C; ++c)
        f.begin[c + 1] += f.begin[c];
      f.to.reserve(pairs.size());
      unsigned selfLoops = 0;
      for (auto &pr : pairs) {
        f.to.push_back(pr.second);
        selfLoops += (pr.second == id);
      }
      f.persistent = (selfLoops == C);
    }
    return f;
  };

  // --- interned cores
  struct CoreStep {
    uint32_t core = 0;           // persistent targets, as a core id
    std::vector<NodeId> loose;   // the other targets, sorted
  };
  struct Core {
    std::vector<NodeId> nodes;   // sorted
    bool stepped = false;
    int32_t mark = 0, plain = 0;
    std::vector<CoreStep> step;  // per class
  };
  std::vector<Core> cores;
  std::unordered_map<std::vector<NodeId>, uint32_t, VecHash> coreIds;
  auto internCore = [&](std::vector<NodeId> &&nodes) -> uint32_t {
    auto it = coreIds.find(nodes);
    if (it != coreIds.end())
      return it->second;
    uint32_t id = static_cast<uint32_t>(cores.size());
    coreIds.emplace(nodes, id);
    cores.emplace_back();
    cores.back().nodes = std::move(nodes);
    return id;
  };
  internCore({});  // core 0: no persistent members

  // splits a sorted node list into persistent and other members
  auto split = [&](const std::vector<NodeId> &all, std::vector<NodeId> &pers,
                   std::vector<NodeId> &loose) {
    pers.clear();
    loose.clear();
    for (NodeId id : all)
      (fanout(id).persistent ? pers : loose).push_back(id);
  };
  std::vector<NodeId> tmpA, tmpB, tmpC;
  auto stepCore = [&](uint32_t coreId) {
    if (cores[coreId].stepped)
      return;
    std::vector<NodeId> members = cores[coreId].nodes;  // cores may grow below
    int32_t mark = 0, plain = 0;
    for (NodeId id : members) {
      const Fanout &f = fanout(id);
      if (f.mark)
        mark = mark ? std::min(mark, f.mark) : f.mark;
      if (nfa[id].result > 0)
        plain = plain ? std::min(plain, nfa[id].result) : nfa[id].result;
    }
    std::vector<CoreStep> step(C);
    for (unsigned c = 0; c < C; ++c) {
      tmpA.clear();
      for (NodeId id : members) {
        const Fanout &f = fan[id];
        tmpA.insert(tmpA.end(), f.to.begin() + f.begin[c], f.to.begin() + f.begin[c + 1]);
      }
      std::sort(tmpA.begin(), tmpA.end());
      tmpA.erase(std::unique(tmpA.begin(), tmpA.end()), tmpA.end());
      split(tmpA, tmpB, step[c].loose);
      step[c].core = internCore(std::move(tmpB));
      tmpB = std::vector<NodeId>();
    }
    Core &core = cores[coreId];
    core.mark = mark;
    core.plain = plain;
    core.step = std::move(step);
    core.stepped = true;
  };

  // --- subsets, keyed by (core id, rest)
  struct Subset {
    uint32_t core;
    std::vector<NodeId> rest;
  };
  std::vector<Subset> subsets;
  std::unordered_map<std::vector<NodeId>, uint32_t, VecHash> ids;  // key: rest + core id
  std::vector<NodeId> key;
  auto intern = [&](uint32_t core, const std::vector<NodeId> &rest) -> uint32_t {
    key = rest;
    key.push_back(static_cast<NodeId>(core));
    auto it = ids.find(key);
    if (it != ids.end())
      return it->second;
    uint32_t id = static_cast<uint32_t>(subsets.size());
    if (nfa.statesUsed() + id + 1 > budget)
      fail_limit("state budget exceeded");
    ids.emplace(key, id);
    subsets.push_back(Subset{core, rest});
    return id;
  };
  intern(0, {});  // 0: the error state
  {
    std::vector<NodeId> init{nfa.initial()}, pers, loose;
    split(init, pers, loose);
    intern(internCore(std::move(pers)), loose);  // 1: the initial state
  }

  std::vector<NodeId> R, Rp, Rl, newRest, merged, lastRest;
  std::vector<uint32_t> rowC(C);
  for (uint32_t cur = 1; cur < subsets.size(); ++cur) {
    const uint32_t coreId = subsets[cur].core;
    const std::vector<NodeId> rest = subsets[cur].rest;  // copy: intern() may reallocate
    stepCore(coreId);
    int32_t result = cores[coreId].mark, plain = cores[coreId].plain;
    for (NodeId id : rest) {
      const Fanout &f = fanout(id);
      if (f.mark)
        result = result ? std::min(result, f.mark) : f.mark;
      if (nfa[id].result > 0)
        plain = plain ? std::min(plain, nfa[id].result) : nfa[id].result;
    }
    dfa.result.resize(subsets.size(), 0);
    dfa.result[cur] = result ? result : plain;

    uint32_t lastCore = 0, lastId = 0;
    bool haveLast = false;
    for (unsigned c = 0; c < C; ++c) {
      R.clear();
      for (NodeId id : rest) {
        const Fanout &f = fan[id];
        R.insert(R.end(), f.to.begin() + f.begin[c], f.to.begin() + f.begin[c + 1]);
      }
      uint32_t newCore = cores[coreId].step[c].core;
      const std::vector<NodeId> *looseOfCore = &cores[coreId].step[c].loose;
      const std::vector<NodeId> *nr = looseOfCore;
      if (!R.empty()) {
        std::sort(R.begin(), R.end());
        R.erase(std::unique(R.begin(), R.end()), R.end());
        split(R, Rp, Rl);
        if (!Rp.empty()) {
          const std::vector<NodeId> &have = cores[newCore].nodes;
          if (!std::includes(have.begin(), have.end(), Rp.begin(), Rp.end())) {
            merged.clear();
            std::set_union(have.begin(), have.end(), Rp.begin(), Rp.end(),
                           std::back_inserter(merged));
            newCore = internCore(std::move(merged));
            merged = std::vector<NodeId>();
            looseOfCore = &cores[coreId].step[c].loose;  // cores may have moved
            nr = looseOfCore;
          }
        }
        if (!Rl.empty()) {
          newRest.clear();
          std::set_union(looseOfCore->begin(), looseOfCore->end(), Rl.begin(), Rl.end(),
                         std::back_inserter(newRest));
          nr = &newRest;
        }
      }
      if (haveLast && (newCore == lastCore) && (*nr == lastRest)) {
        rowC[c] = lastId;
        continue;
      }
      lastRest = *nr;
      lastCore = newCore;
      lastId = intern(newCore, lastRest);
      haveLast = true;
      rowC[c] = lastId;
    }
    dfa.next.resize(subsets.size() * 256, 0);
    for (unsigned b = 0; b < 256; ++b)
      dfa.next[static_cast<size_t>(cur) * 256 + b] = rowC[clsOf[b]];
  }
  dfa.n = static_cast<uint32_t>(subsets.size());
  if (getenv("RED_B200_COMPILE_TIMING")) {
    size_t tot = 0;
    for (auto &v : subsets)
      tot += v.rest.size();
    fprintf(stderr, "red_b200 determinize: nfa %zu classes %u subsets %zu cores %zu rest members %zu\n",
            nNodes, C, subsets.size(), cores.size(), tot);
  }
  dfa.result.resize(dfa.n, 0);
  dfa.next.resize(static_cast<size_t>(dfa.n) * 256, 0);
  return dfa;
}

// byte -> class; classes numbered by first occurrence
std::vector<uint8_t> byte_classes(const RawDfa &dfa, uint32_t &classCount) {
  std::map<std::vector<uint32_t>, uint32_t> seen;
  std::vector<uint8_t> cls(256);
  std::vector<uint32_t> col(dfa.n);
  for (unsigned c = 0; c < 256; ++c) {
    for (uint32_t s = 0; s < dfa.n; ++s)
      col[s] = dfa.next[static_cast<size_t>(s) * 256 + c];
    auto it = seen.find(col);
    if (it == seen.end())
      it = seen.emplace(col, static_cast<uint32_t>(seen.size())).first;
    cls[c] = static_cast<uint8_t>(it->second);
  }
  classCount = static_cast<uint32_t>(seen.size());
  return cls;
}

/* The reference's minimizer (Gries' description of Hopcroft's algorithm,
   lib/Minimizer.cpp:104-300) seeds its work list with EITHER the rejecting
   block OR all accepting blocks, whichever side has fewer states
   (makeList, Minimizer.cpp:189).  With one accepting block that is the
   classic "all but one" and exact.  With several result ids and at least as
   many accepting as rejecting states, the accepting blocks are never used as
   splitters unless they get split themselves, and states that differ only in
   WHICH result they lead to stay merged: for `.*ab.*`->1, `.*x.*`->2 the
   reference's program answers 2 for "xab" although its own unminimized DFA
   says 1.  Programs in the field were built with that behaviour, so the
   default here follows the same rule - same seeding, same "both halves if
   pending, else the smaller half" patching, splits applied in increasing
   state order - and kMinimizeExact gives the true minimal DFA instead.
   The one thing not reproduced is the order in which the reference's hash
   set hands out work items, which its result can depend on in contrived
   cases; tests/test_compile.py measures the agreement. */
static size_t refine_like_reference(const RawDfa &dfa,
                                    const std::vector<unsigned> &rep, uint32_t C,
                                    uint32_t nInitial,
                                    std::vector<uint32_t> &block) {
  const uint32_t n = dfa.n;
  auto delta = [&](uint32_t s, uint32_t c) {
    return dfa.next[static_cast<size_t>(s) * 256 + rep[c]];
  };
  // inverse transitions: pre[c][t] = states s with delta(s, c) == t
  std::vector<std::vector<std::vector<uint32_t>>> pre(C, std::vector<std::vector<uint32_t>>(n));
  for (uint32_t s = 0; s < n; ++s)
    for (uint32_t c = 0; c < C; ++c)
      pre[c][delta(s, c)].push_back(s);
  std::vector<std::vector<uint32_t>> members(nInitial);
  for (uint32_t s = 0; s < n; ++s)
    members[block[s]].push_back(s);

  std::set<std::pair<uint32_t, uint32_t>> list;  // (block, class)
  {
    size_t zero = members[0].size(), rest = 0;
    for (uint32_t b = 1; b < nInitial; ++b)
      rest += members[b].size();
    uint32_t lo = 0, hi = 1;
    if ((rest > 0) && (rest < zero)) {
      lo = 1;
      hi = nInitial;
    }
    for (uint32_t b = lo; b < hi; ++b)
      for (uint32_t c = 0; c < C; ++c)
        list.emplace(b, c);
  }

  std::vector<uint32_t> mark(n, 0);
  uint32_t stamp = 0;
  while (!list.empty()) {
    auto item = *list.begin();
    list.erase(list.begin());
    const uint32_t B = item.first, c = item.second;
    // states with a c-transition into B, ascending
    std::vector<uint32_t> splits;
    ++stamp;
    for (uint32_t t : members[B])
      for (uint32_t s : pre[c][t])
        if (mark[s] != stamp) {
          mark[s] = stamp;
          splits.push_back(s);
        }
    if (splits.empty())
      continue;
    std::sort(splits.begin(), splits.end());
    std::map<uint32_t, uint32_t> twin;  // block -> its new sibling
    std::vector<std::pair<uint32_t, uint32_t>> patches;
    for (uint32_t s : splits) {
      const uint32_t X = block[s];
      // does every state still in X go into (what is still in) B on c?
      bool contained = true;
      for (uint32_t m : members[X])
        if (block[delta(m, c)] != B) {
          contained = false;
          break;
        }
      if (contained)
        continue;
      auto it = twin.find(X);
      if (it == twin.end()) {
        it = twin.emplace(X, static_cast<uint32_t>(members.size())).first;
        members.emplace_back();
        patches.emplace_back(X, it->second);
      }
      std::vector<uint32_t> &mx = members[X];
      mx.erase(std::find(mx.begin(), mx.end(), s));
      members[it->second].push_back(s);
      block[s] = it->second;
    }
    for (auto &p : patches) {
      const uint32_t ii = p.first, jj = p.second;
      const uint32_t small = (members[ii].size() < members[jj].size()) ? ii : jj;
      for (uint32_t ch = 0; ch < C; ++ch) {
        if (list.count({ii, ch}))
          list.emplace(jj, ch);
        else
          list.emplace(small, ch);
      }
    }
  }
  // compact block numbers (a block can have been emptied)
  std::vector<uint32_t> renum(members.size(), 0xffffffffu);
  uint32_t k = 0;
  for (uint32_t s = 0; s < n; ++s) {
    if (renum[block[s]] == 0xffffffffu)
      renum[block[s]] = k++;
    block[s] = renum[block[s]];
  }
  return k;
}

MinDfa minimize(const RawDfa &dfa, MinimizeMode mode) {
  MinDfa out;
  out.equiv = byte_classes(dfa, out.classCount);
  const uint32_t C = out.classCount;
  const uint32_t n = dfa.n;
  std::vector<unsigned> rep(C);  // a representative byte per class
  for (int c = 255; c >= 0; --c)
    rep[out.equiv[c]] = static_cast<unsigned>(c);

  // block[s] starts as the rank of result(s) (reference: partition(),
  // lib/Minimizer.cpp:166)
  std::vector<uint32_t> block(n);
  uint32_t nInitial = 0;
  {
    std::map<int32_t, uint32_t> rank;
    rank[0] = 0;
    for (uint32_t s = 0; s < n; ++s)
      rank.emplace(dfa.result[s], 0);
    for (auto &kv : rank)
      kv.second = nInitial++;
    for (uint32_t s = 0; s < n; ++s)
      block[s] = rank[dfa.result[s]];
  }
  size_t blocks = 0;
  if (mode == kMinimizeExact) {
    // Moore refinement to the true minimal DFA
    for (;;) {
      std::unordered_map<std::string, uint32_t> index;
      std::vector<uint32_t> nextBlock(n);
      std::string key;
      key.resize((C + 1) * 4);
      for (uint32_t s = 0; s < n; ++s) {
        uint32_t *k = reinterpret_cast<uint32_t *>(&key[0]);
        k[0] = block[s];
        for (uint32_t c = 0; c < C; ++c)
          k[1 + c] = block[dfa.next[static_cast<size_t>(s) * 256 + rep[c]]];
        auto it = index.find(key);
        if (it == index.end())
          it = index.emplace(key, static_cast<uint32_t>(index.size())).first;
        nextBlock[s] = it->second;
      }
      block.swap(nextBlock);
      if (index.size() == blocks)
        break;
      blocks = index.size();
    }
  }
  else {
    blocks = refine_like_reference(dfa, rep, C, nInitial, block);
  }

  // number the blocks: error 0, initial 1, the rest breadth-first by class.
  // A block is represented by its earliest-created state, as in the reference
  // (makeDfaFromBlocks, lib/Minimizer.cpp:318: the first member of the block);
  // this only matters when the reference-like refinement left a block whose
  // members are not all equivalent.
  const uint32_t none = 0xffffffffu;
  std::vector<uint32_t> firstOf(blocks, none);
  for (uint32_t s = 0; s < n; ++s)
    if (firstOf[block[s]] == none)
      firstOf[block[s]] = s;
  std::vector<uint32_t> blockToNew(blocks, none);
  std::vector<uint32_t> repState;  // new id -> the state whose row it takes
  blockToNew[block[0]] = 0;
  repState.push_back(firstOf[block[0]]);
  const bool initIsError = block[1] == block[0];
  if (!initIsError)
    blockToNew[block[1]] = 1;
  repState.push_back(initIsError ? 1 : firstOf[block[1]]);
  // (if the initial state shares the error state's block it still gets its
  //  own number, with every transition leading to the error state)
  for (uint32_t id = 1; id < repState.size(); ++id) {
    if ((id == 1) && initIsError)
      continue;
    uint32_t s = repState[id];
    for (uint32_t c = 0; c < C; ++c) {
      uint32_t b = block[dfa.next[static_cast<size_t>(s) * 256 + rep[c]]];
      if (blockToNew[b] == none) {
        blockToNew[b] = static_cast<uint32_t>(repState.size());
        repState.push_back(firstOf[b]);
      }
    }
  }
  // Blocks that the representatives' rows no longer reach still get a state,
  // as in the reference (one state per block): the state count is part of the
  // file and of the choice of value width.
  for (uint32_t b = 0; b < blocks; ++b)
    if (blockToNew[b] == none) {
      blockToNew[b] = static_cast<uint32_t>(repState.size());
      repState.push_back(firstOf[b]);
    }
  out.n = static_cast<uint32_t>(repState.size());
  out.next.assign(static_cast<size_t>(out.n) * C, 0);
  out.result.assign(out.n, 0);
  for (uint32_t id = 0; id < out.n; ++id) {
    if ((id == 1) && initIsError)
      continue;
    uint32_t s = repState[id];
    out.result[id] = dfa.result[s];
    for (uint32_t c = 0; c < C; ++c)
      out.next[static_cast<size_t>(id) * C + c] =
          blockToNew[block[dfa.next[static_cast<size_t>(s) * 256 + rep[c]]]];
  }

  // dead-end flag exactly as the reference computes it (lib/Dfa.cpp:30):
  // every non-error transition loops back, and - except for state 0 - at
  // least maxChar of the maxChar+1 classes do
  const uint32_t maxChar = C - 1;
  out.deadEnd.assign(out.n, 0);
  for (uint32_t id = 0; id < out.n; ++id) {
    uint32_t nonError = 0;
    bool loops = true;
    for (uint32_t c = 0; c < C; ++c) {
      uint32_t t = out.next[static_cast<size_t>(id) * C + c];
      if (t != 0) {
        ++nonError;
        if (t != id)
          loops = false;
      }
    }
    bool dead = loops;
    if (dead && (nonError < maxChar) && (id != 0))
      dead = false;
    out.deadEnd[id] = dead;
  }

  // required prefix (lib/Dfa.cpp:166): follow states that reject and have
  // exactly one non-error class
  out.leaderNext = 1;
  uint32_t id = 1;
  for (int i = 0; i < 256; ++i) {
    if (out.result[id] > 0)
      break;
    uint32_t only = 0, cnt = 0, cls = 0;
    for (uint32_t c = 0; c < C; ++c) {
      uint32_t t = out.next[static_cast<size_t>(id) * C + c];
      if (t != 0) {
        ++cnt;
        only = t;
        cls = c;
      }
    }
    if (cnt != 1)
      break;
    out.leader.push_back(static_cast<uint8_t>(cls));
    out.leaderNext = only;
    if (only == id)
      break;
    id = only;
  }
  if (out.leader.size() > 255)
    fail_internal("leader too long");
  return out;
}

namespace {

struct FmtLimits { int64_t maxResult; uint64_t maxOffset; };

FmtLimits limits(int fmt) {
  switch (fmt) {
  case 1: return {0x7f, 0xff};
  case 2: return {0x7fff, 0xffff};
  default: return {0x7fffffff, 0xffffffffull};
  }
}

bool resultFits(int fmt, int32_t r) { return r <= limits(fmt).maxResult; }
bool offsetFits(int fmt, uint32_t maxChar, uint64_t n) {
  // same sum as the reference (Proxy.h:161): rows of maxChar + 2 values
  return n * maxChar + n + n <= limits(fmt).maxOffset;
}

void put(std::string &buf, uint64_t v, int width) {
  buf.append(reinterpret_cast<const char *>(&v), width);  // little endian host
}

} // anonymous

std::string serialize(const MinDfa &dfa, int fmt) {
  const uint32_t maxChar = dfa.classCount - 1;
  int32_t maxResult = 0;
  for (int32_t r : dfa.result)
    maxResult = std::max(maxResult, r);

  if (fmt == RED_B200_FMT_AUTO) {
    // smallest width that holds both the results and the offsets
    int byResult = !resultFits(2, maxResult) ? 4 : !resultFits(1, maxResult) ? 2 : 1;
    if (!offsetFits(4, maxChar, dfa.n))
      fail_limit("num states too many for any format");
    int byOffset = !offsetFits(2, maxChar, dfa.n) ? 4 : !offsetFits(1, maxChar, dfa.n) ? 2 : 1;
    fmt = std::max(byResult, byOffset);
  }
  if ((fmt != 1) && (fmt != 2) && (fmt != 4))
    fail_internal("unsuitable dfa format requested");
  if (!resultFits(fmt, maxResult) || !offsetFits(fmt, maxChar, dfa.n))
    fail_limit("dfa too big for format");

  const uint64_t rowVals = static_cast<uint64_t>(maxChar) + 2;
  const uint64_t rowBytes = rowVals * fmt;
  if (rowBytes * dfa.n > 0xffffffffull)
    fail_internal("initial offset too large");

  std::string buf;
  buf.reserve(kHeaderSize + 256 + rowBytes * dfa.n);
  buf.append("REDA", 4);
  put(buf, 1, 2);                 // major
  put(buf, 0, 2);                 // minor
  put(buf, 0, 4);                 // checksum, patched below
  put(buf, fmt, 1);
  put(buf, maxChar, 1);
  put(buf, dfa.leader.size(), 1);
  put(buf, 0, 1);
  put(buf, dfa.n, 4);
  put(buf, rowBytes * 1, 4);                 // initialOff_: byte offset of state 1
  put(buf, rowBytes * dfa.leaderNext, 4);    // leaderOff_
  put(buf, 0, 4);
  buf.append(reinterpret_cast<const char *>(dfa.equiv.data()), 256);
  buf.append(reinterpret_cast<const char *>(dfa.leader.data()), dfa.leader.size());
  buf.append(((dfa.leader.size() + 7) & ~size_t(7)) - dfa.leader.size(), '\0');

  const uint64_t topBit = 8ull * fmt - 1;
  const uint64_t mask = (1ull << topBit) - 1;
  for (uint32_t id = 0; id < dfa.n; ++id) {
    uint64_t head = (static_cast<uint64_t>(dfa.result[id]) & mask) |
                    (static_cast<uint64_t>(dfa.deadEnd[id]) << topBit);
    put(buf, head, fmt);
    for (uint32_t c = 0; c <= maxChar; ++c)
      put(buf, dfa.next[static_cast<size_t>(id) * dfa.classCount + c] * rowVals, fmt);
  }
  uint32_t sum = calc_checksum(buf.data(), buf.size());
  memcpy(&buf[8], &sum, 4);
  return buf;
}

} // namespace rx
} // namespace red_b200
