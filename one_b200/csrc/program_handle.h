/* program_handle.h - the object behind `red_b200_program *` (private). */

#pragma once

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/red_b200.h"
#include "device_program.h"
#include "hot.h"
#include "program.h"

struct HostPipe;
struct GrepPipe;
struct SinglePipe;

using red_b200::BatchArgs;
using red_b200::BatchTuning;
using red_b200::DeviceImage;
using red_b200::LaunchGeometry;
using red_b200::ParsedProgram;
using red_b200::ProgramDev;


struct red_b200_program {
  int device = 0;
  LaunchGeometry geo{};
  DeviceImage img;      // host copy (tables cleared after upload)
  ParsedProgram parsed; // kept small: tables cleared after upload
  ProgramDev dev{};
  bool smemTable = false;
  uint32_t format = 0;
  uint32_t leaderLen = 0;
  uint64_t tableBytes = 0;
  void *dTable = nullptr;
  void *dCls = nullptr;
  void *dResults = nullptr;
  void *dLeader = nullptr;
  BatchTuning tune;
  std::mutex mu;        // guards the lazily created host pipeline
  HostPipe *pipe = nullptr;
  GrepPipe *grep = nullptr;   // lazily created by red_b200_grep_host
  SinglePipe *single = nullptr;  // lazily created by red_b200_match on a short text
  std::atomic<long> launches{0};  // kernels launched through this program (for bench)
  std::atomic<long> hotLaunches{0};  // of those: launches of the staged batch kernel

  // the hot image of the staged batch kernel (kernels_hot.cu)
  red_b200::HotDev hotDev[2]{};        // [0] style Instant, [1] styles Last and Full
  red_b200::ProgramDev devBatch{};     // global-table programs: the same DFA with the states a profiled
                                       // walk enters most in front (the shared-memory mirror's rows);
                                       // check/match batches only - states handed to callers stay prog->dev's
  bool haveDevBatch = false;
  size_t smemBudget = 0;
  uint32_t hotStreamStates = 0;        // cap on that image's states (0: what fits)
  red_b200::HotDev hotStream{};        // the long-stream matcher's (two-byte codes where a byte is too small)
  red_b200::HotTuning hotTune;
  bool hotEnabled = true;
  uint32_t hotMaxStates = red_b200::kHotMax;
  std::atomic<int> hotProfiled{0};    // 1: the hot set was chosen from measured visit counts
  bool hotProfileWanted = true;
  int hotAsciiRows = 0;               // rows of 128 columns: 0 = by the profile, 1 = always, -1 = never
  std::atomic<uint32_t> avgLineBytes{0};  // from the profile sample (0: unknown)
  std::mutex hotMu;                   // guards re-profiling
  std::vector<void *> hotBuffers;     // every device buffer a hot image ever used (freed with the program)
};


namespace red_b200 {
// launch the per-line kernels for one batch (capi.cu)
int launch_batch_for(red_b200_program *prog, const BatchArgs &args, int verb,
                     int style, cudaStream_t stream, std::string &why);
void release_grep_pipe(red_b200_program *prog);  // grep.cu
void set_last_error(const std::string &msg);      // capi.cu: what red_b200_last_error() returns
}
