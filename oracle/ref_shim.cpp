/* ref_shim.cpp - C-ABI shim around the UNMODIFIED reference (zezax::red).

   TEST INFRASTRUCTURE ONLY.  This file is compiled together with the
   reference's own lib/ sources, in place, from /root/reference/quol/red
   (see oracle/Makefile) into oracle/_ref/libred_ref.so.  Nothing under
   one_b200/ links or loads it.  It is used by
     - tests/ (as the checker and to regenerate tests/golden/),
     - bench.py's cpu_baseline leg and `--impl reference` arm.

   Every entry point is a thin call into the reference's public API:
     Parser::addAs            (include/Parser.h:98)
     compileToSerialized      (include/Compile.h:30)
     check/match/scan/search  (include/Matcher.h:74-94)
     replace / matchAll       (include/Matcher.h:96-129)
     Red::collect             (include/Red.h:109)
     StatefulMatcher          (include/Matcher.h:770-794)
 */

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "Compile.h"
#include "Executable.h"
#include "Matcher.h"
#include "Parser.h"
#include "Red.h"

using namespace zezax::red;

namespace {

thread_local std::string g_err;

int fail(const std::exception &ex) {
  g_err = ex.what();
  return -1;
}

struct RefProg {
  Executable exec;
  explicit RefProg(std::string_view sv) : exec(gCopyTag, sv) {}
};

enum Verb { vCheck = 0, vMatch = 1, vScan = 2, vSearch = 3 };

inline void runOne(const Executable &ex, const char *p, size_t n, int verb,
                   Style sty, int32_t &res, uint64_t &st, uint64_t &en) {
  switch (verb) {
  case vCheck:
    res = check(ex, p, n, sty); st = 0; en = 0; break;
  case vScan:
    res = scan(ex, p, n, sty); st = 0; en = 0; break;
  case vMatch: {
    Outcome oc = match(ex, p, n, sty);
    res = oc.result_; st = oc.start_; en = oc.end_; break; }
  default: {
    Outcome oc = search(ex, p, n, sty);
    res = oc.result_; st = oc.start_; en = oc.end_; break; }
  }
}

} // anonymous

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

/* Compile n patterns into one serialized DFA (malloc'd; caller ref_free()s).
   langs[i]: 1 raw regex, 2 auto regex, 3 glob, 4 exact (Parser.h:43-49).
   fmt: 1/2/4 or 255 for auto (Serializer.h:37-43).
   max_states: 0 = no Budget, else Budget(states) (Budget.h:33). */
int ref_compile(int n, const char *const *pats, const uint64_t *patLens,
                const int32_t *results, const uint32_t *flags,
                const int32_t *langs, int fmt, uint64_t maxStates,
                char **outBuf, uint64_t *outLen) {
  try {
    Budget budget;
    if (maxStates)
      budget.initStates(maxStates);
    Parser p(maxStates ? &budget : nullptr);
    for (int ii = 0; ii < n; ++ii)
      p.addAs(static_cast<Language>(langs[ii]),
              std::string_view(pats[ii], patLens[ii]), results[ii], flags[ii]);
    std::string buf = compileToSerialized(p, static_cast<Format>(fmt));
    char *mem = static_cast<char *>(malloc(buf.size()));
    memcpy(mem, buf.data(), buf.size());
    *outBuf = mem;
    *outLen = buf.size();
    return 0;
  }
  catch (const RedExceptLimit &ex) { g_err = ex.what(); return -2; }
  catch (const RedExceptParse &ex) { g_err = ex.what(); return -3; }
  catch (const RedExceptApi &ex) { g_err = ex.what(); return -4; }
  catch (const std::exception &ex) { return fail(ex); }
}

void ref_free(void *p) { free(p); }

void *ref_load(const char *prog, uint64_t len) {
  try {
    return new RefProg(std::string_view(prog, len));
  }
  catch (const std::exception &ex) { fail(ex); return nullptr; }
}

void ref_unload(void *h) { delete static_cast<RefProg *>(h); }

/* verb: 0 check, 1 match, 2 scan, 3 search; style: Matcher.h:62-69 */
int ref_run(void *h, const char *text, uint64_t len, int verb, int style,
            int32_t *res, uint64_t *start, uint64_t *end) {
  try {
    runOne(static_cast<RefProg *>(h)->exec, text, len, verb,
           static_cast<Style>(style), *res, *start, *end);
    return 0;
  }
  catch (const std::exception &ex) { return fail(ex); }
}

/* The reference has no batch call; this is the loop a RED user writes:
   one match()/check()/... per line, lines spread over nthreads threads
   (doc/Performance.md "Threads": matching is read-only and lock-free).
   Lines are text[starts[i] .. starts[i]+lens[i]).  starts_out/ends may be
   null (then only results are stored). */
int ref_run_batch(void *h, const char *text, const uint64_t *starts,
                  const uint32_t *lens, uint64_t n, int verb, int style,
                  int nthreads, int32_t *results, uint32_t *startsOut,
                  uint32_t *endsOut) {
  try {
    const Executable &ex = static_cast<RefProg *>(h)->exec;
    Style sty = static_cast<Style>(style);
    if ((sty < styInstant) || (sty > styFull))
      throw RedExceptExec("unsupported style");
    if (nthreads < 1)
      nthreads = 1;
    auto work = [&](uint64_t lo, uint64_t hi) {
      for (uint64_t ii = lo; ii < hi; ++ii) {
        int32_t r; uint64_t s, e;
        runOne(ex, text + starts[ii], lens[ii], verb, sty, r, s, e);
        results[ii] = r;
        if (startsOut) startsOut[ii] = static_cast<uint32_t>(s);
        if (endsOut) endsOut[ii] = static_cast<uint32_t>(e);
      }
    };
    if (nthreads == 1) {
      work(0, n);
      return 0;
    }
    std::vector<std::thread> pool;
    uint64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
      uint64_t lo = std::min<uint64_t>(n, per * t);
      uint64_t hi = std::min<uint64_t>(n, lo + per);
      if (lo < hi)
        pool.emplace_back(work, lo, hi);
    }
    for (std::thread &th : pool)
      th.join();
    return 0;
  }
  catch (const std::exception &ex) { return fail(ex); }
}

/* replace(): Matcher.h:96-102.  Output malloc'd. */
int ref_replace(void *h, const char *text, uint64_t len, const char *repl,
                uint64_t replLen, uint64_t max, int style, char **outBuf,
                uint64_t *outLen, uint64_t *count) {
  try {
    std::string out;
    *count = replace(static_cast<RefProg *>(h)->exec, text, len,
                     std::string_view(repl, replLen), out, max,
                     static_cast<Style>(style));
    char *mem = static_cast<char *>(malloc(out.size() + 1));
    memcpy(mem, out.data(), out.size());
    *outBuf = mem;
    *outLen = out.size();
    return 0;
  }
  catch (const std::exception &ex) { return fail(ex); }
}

/* A fresh StatefulMatcher fed the text byte by byte: results[0] is result()
   before any input, results[1 + i] what advance(text[i]) returned. */
int ref_stateful(void *h, const char *text, uint64_t len, int32_t *results) {
  try {
    StatefulMatcher sm(static_cast<RefProg *>(h)->exec);
    results[0] = sm.result();
    for (uint64_t ii = 0; ii < len; ++ii)
      results[1 + ii] = sm.advance(static_cast<Byte>(text[ii]));
    return 0;
  }
  catch (const std::exception &ex) { return fail(ex); }
}

/* which: 0 = matchAll() (Matcher.h:127), 1 = Red::collect() (Red.h:109).
   Fills up to cap outcomes; returns total count in *count. */
int ref_multi(void *h, const char *text, uint64_t len, int which,
              uint64_t cap, int32_t *results, uint64_t *starts,
              uint64_t *ends, uint64_t *count) {
  try {
    std::vector<Outcome> out;
    RefProg *rp = static_cast<RefProg *>(h);
    if (which == 0)
      matchAll(rp->exec, std::string_view(text, len), out);
    else {
      Red re(gUnownedTag, rp->exec.serialized());
      re.collect(std::string_view(text, len), out);
    }
    *count = out.size();
    for (size_t ii = 0; (ii < out.size()) && (ii < cap); ++ii) {
      results[ii] = out[ii].result_;
      starts[ii] = out[ii].start_;
      ends[ii] = out[ii].end_;
    }
    return 0;
  }
  catch (const std::exception &ex) { return fail(ex); }
}

} // extern "C"
