/* stream_match.cu - one long byte stream (placeholder until the chunked
   matcher lands in this file). */
#include "stream.h"

namespace red_b200 {

int match_stream(const ProgramDev &, bool, const DeviceImage &,
                 const LaunchGeometry &, const uint8_t *, uint64_t, int,
                 red_b200_outcome *, cudaStream_t, int *, std::string &why) {
  why = "stream matcher not built yet";
  return RED_B200_EINTERNAL;
}

} // namespace red_b200
