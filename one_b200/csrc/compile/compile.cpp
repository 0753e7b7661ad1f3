/* compile.cpp - patterns -> serialized REDA program: the host-side compile
   pipeline behind red_b200_compile() (reference: compileToSerialized,
   lib/Compile.cpp:26). */

#include "compile.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "../../../include/red_b200.h"
#include "dfa.h"
#include "parser.h"

namespace red_b200 {

int compile_patterns(int n, const char *const *patterns,
                     const uint64_t *patternLens, const int32_t *results,
                     const uint32_t *flags, const int32_t *langs, int format,
                     uint64_t maxStates, std::string &out, std::string &why) {
  try {
    const uint64_t budget = maxStates ? maxStates : ~0ull;
    rx::PatternParser parser(maxStates, 0);
    for (int i = 0; i < n; ++i)
      parser.add(langs[i], reinterpret_cast<const uint8_t *>(patterns[i]),
                 patternLens[i], results[i], flags[i]);
    const bool timing = getenv("RED_B200_COMPILE_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
      if (!timing)
        return;
      auto t1 = std::chrono::steady_clock::now();
      fprintf(stderr, "red_b200 compile: %s %.3f s\n", what,
              std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
    };
    lap("parse");
    rx::Nfa &nfa = parser.finish();
    lap("finish");
    rx::RawDfa raw = rx::determinize(nfa, budget);
    lap("determinize");
    // The true minimal DFA by default: it accepts exactly what the rules say.
    // RED_B200_MINIMIZE=reference follows the reference's work-list rule
    // instead, which under-splits some multi-result automata (see dfa.cpp) -
    // for comparing programs state by state with the reference's.
    const char *env = getenv("RED_B200_MINIMIZE");
    rx::MinimizeMode mode = (env && (std::string(env) == "reference"))
        ? rx::kMinimizeLikeReference : rx::kMinimizeExact;
    rx::MinDfa min = rx::minimize(raw, mode);
    lap("minimize");
    out = rx::serialize(min, format);
    lap("serialize");
    return RED_B200_OK;
  }
  catch (const rx::CompileError &ex) {
    why = ex.msg;
    return ex.code;
  }
  catch (const std::bad_alloc &) {
    why = "out of memory";
    return RED_B200_ELIMIT;
  }
}

} // namespace red_b200
