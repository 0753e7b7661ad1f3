/* dfa.h - determinize, minimize, serialize (see dfa.cpp). */

#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "nfa.h"

namespace red_b200 {
namespace rx {

// unminimized DFA over raw bytes; state 0 = error, state 1 = initial
struct RawDfa {
  uint32_t n = 0;
  std::vector<uint32_t> next;   // n * 256
  std::vector<int32_t> result;  // n
};

// minimal DFA over byte classes; state 0 = error, state 1 = initial
struct MinDfa {
  uint32_t n = 0;
  uint32_t classCount = 0;
  std::vector<uint8_t> equiv;    // 256: byte -> class
  std::vector<uint32_t> next;    // n * classCount
  std::vector<int32_t> result;   // n
  std::vector<uint8_t> deadEnd;  // n
  std::vector<uint8_t> leader;   // classes of the required prefix
  uint32_t leaderNext = 1;       // state reached after the leader
};

enum MinimizeMode {
  kMinimizeLikeReference = 0,  // the reference's work-list rule (see dfa.cpp)
  kMinimizeExact = 1           // true minimal DFA
};

RawDfa determinize(const Nfa &nfa, uint64_t stateBudget);
MinDfa minimize(const RawDfa &dfa, MinimizeMode mode);
std::string serialize(const MinDfa &dfa, int format);

} // namespace rx
} // namespace red_b200
