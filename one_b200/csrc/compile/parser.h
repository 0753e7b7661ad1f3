/* parser.h - pattern languages -> NFA (see parser.cpp). */

#pragma once

#include <cstddef>
#include <cstdint>

#include "nfa.h"

namespace red_b200 {
namespace rx {

class PatternParser {
public:
  // maxStates / maxParens: 0 = unlimited (reference: Budget.h)
  PatternParser(uint64_t maxStates, uint64_t maxParens);
  ~PatternParser();
  PatternParser(const PatternParser &) = delete;
  PatternParser &operator=(const PatternParser &) = delete;

  // lang: RED_B200_LANG_*; flags: RED_B200_F_*; throws CompileError
  void add(int lang, const uint8_t *pattern, size_t len, int32_t result,
           uint32_t flags);

  // no more patterns: returns the finished automaton
  Nfa &finish();

private:
  struct Impl;
  Impl *impl_;
};

} // namespace rx
} // namespace red_b200
