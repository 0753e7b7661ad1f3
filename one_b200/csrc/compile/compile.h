/* compile.h - regex/glob/exact patterns -> serialized REDA program.
   Host-side C++; the GPU is not involved. */

#pragma once

#include <cstdint>
#include <string>

namespace red_b200 {

// returns a RED_B200_* code; on failure `why` holds the message
int compile_patterns(int n, const char *const *patterns,
                     const uint64_t *patternLens, const int32_t *results,
                     const uint32_t *flags, const int32_t *langs, int format,
                     uint64_t maxStates, std::string &out, std::string &why);

} // namespace red_b200
