/* lanes_dispatch.cuh - runtime (style, positions, table layout) -> template
   instantiation, shared by the translation units that each hold one schedule
   of the batch kernels (kernels_lanes.cu, kernels_tile.cu,
   kernels_sorted.cu - one schedule per file so that they compile in
   parallel).  L::go<STYLE, WANT_START, SMEM, K, ClsT>(...) launches. */

#pragma once

#include "chunk_engine.cuh"

#include <algorithm>

namespace red_b200 {
namespace dev {

struct LaunchShape {
  int threads;
  int perSm;
  size_t smem;  // class table + transition table (or hot rows) + alignment pad
};

template <bool SMEM>
inline LaunchShape launch_shape(const ProgramDev &pd, const BatchTuning &tune) {
  LaunchShape sh;
  sh.threads = tune.threads ? tune.threads : kMaxThreads;
  sh.perSm = tune.ctasPerSm ? tune.ctasPerSm : 1;
  sh.smem = kLanesSmemExtra + (SMEM ? pd.tableBytes : pd.hotBytes);
  return sh;
}

template <class L, int STYLE, bool WANT_START>
cudaError_t dispatch_table(const ProgramDev &pd, bool smemTable,
                           const BatchArgs &args, const LaunchGeometry &geo,
                           const BatchTuning &tune, unsigned long long *counter,
                           cudaStream_t stream) {
  const bool cls8 = pd.clsMaxOffset < 256;
  if (smemTable) {
    if (pd.unitShift == 0)
      return cls8 ? L::template go<STYLE, WANT_START, true, 1, uint8_t>(pd, args, geo, tune, counter, stream)
                  : L::template go<STYLE, WANT_START, true, 1, uint16_t>(pd, args, geo, tune, counter, stream);
    if (pd.unitShift == 1)
      return cls8 ? L::template go<STYLE, WANT_START, true, 2, uint8_t>(pd, args, geo, tune, counter, stream)
                  : L::template go<STYLE, WANT_START, true, 2, uint16_t>(pd, args, geo, tune, counter, stream);
    if (pd.unitShift == 8)  // byte-indexed rows: one load per byte
      return L::template go<STYLE, WANT_START, true, 9, uint16_t>(pd, args, geo, tune, counter, stream);
    return cudaErrorInvalidValue;
  }
  return cls8 ? L::template go<STYLE, WANT_START, false, 0, uint8_t>(pd, args, geo, tune, counter, stream)
              : L::template go<STYLE, WANT_START, false, 0, uint16_t>(pd, args, geo, tune, counter, stream);
}

template <class L, bool WANT_START>
cudaError_t dispatch_style(const ProgramDev &pd, bool smemTable,
                           const BatchArgs &args, int style,
                           const LaunchGeometry &geo, const BatchTuning &tune,
                           unsigned long long *counter, cudaStream_t stream) {
  switch (style) {
  case STY_INSTANT: return dispatch_table<L, STY_INSTANT, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_FIRST:   return dispatch_table<L, STY_FIRST, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_TANGENT: return dispatch_table<L, STY_TANGENT, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_LAST:    return dispatch_table<L, STY_LAST, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  case STY_FULL:    return dispatch_table<L, STY_FULL, WANT_START>(pd, smemTable, args, geo, tune, counter, stream);
  default:          return cudaErrorInvalidValue;
  }
}

// zeroes the work counter, then dispatches; args.matchStarts selects WANT_START
template <class L>
cudaError_t dispatch_family(const ProgramDev &pd, bool smemTable,
                            const BatchArgs &args, int style,
                            const LaunchGeometry &geo, const BatchTuning &tune,
                            unsigned long long *counter, cudaStream_t stream) {
  if (args.nLines == 0)
    return cudaSuccess;
  cudaError_t err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
  if (err != cudaSuccess)
    return err;
  return args.matchStarts
      ? dispatch_style<L, true>(pd, smemTable, args, style, geo, tune, counter, stream)
      : dispatch_style<L, false>(pd, smemTable, args, style, geo, tune, counter, stream);
}

} // namespace dev

// one per translation unit
cudaError_t launch_family_lanes(const ProgramDev &pd, bool smemTable, const BatchArgs &args, int style,
                                const LaunchGeometry &geo, const BatchTuning &tune,
                                unsigned long long *counter, cudaStream_t stream);
cudaError_t launch_family_tile(const ProgramDev &pd, bool smemTable, const BatchArgs &args, int style,
                               const LaunchGeometry &geo, const BatchTuning &tune,
                               unsigned long long *counter, cudaStream_t stream);
// *kernels: kernels launched (pre-sort + batch); 0 = the scratch does not fit
// beside the table, nothing was launched
cudaError_t launch_family_sorted(const ProgramDev &pd, bool smemTable, const BatchArgs &args, int style,
                                 const LaunchGeometry &geo, const BatchTuning &tune,
                                 unsigned long long *counter, cudaStream_t stream, int *kernels);

} // namespace red_b200
