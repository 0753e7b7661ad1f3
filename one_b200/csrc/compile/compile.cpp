/* compile.cpp - placeholder until the host compile pipeline lands. */
#include "compile.h"
#include "../../../include/red_b200.h"

namespace red_b200 {

int compile_patterns(int, const char *const *, const uint64_t *,
                     const int32_t *, const uint32_t *, const int32_t *, int,
                     uint64_t, std::string &, std::string &why) {
  why = "compile pipeline not built yet";
  return RED_B200_EINTERNAL;
}

} // namespace red_b200
