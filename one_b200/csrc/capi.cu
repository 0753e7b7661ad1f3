/* capi.cu - the C ABI of libred_b200.so (include/red_b200.h).

   Host-side plumbing only: argument checks with the reference's error
   classes, program upload, launches, and the chunked host<->device pipeline
   of the *_host entry points.  There is no CPU execution path in this file:
   if no sm_100 device is present the compute entry points fail. */

#include "../../include/red_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "compile/compile.h"
#include "device_program.h"
#include "program.h"
#include "program_handle.h"
#include "stream.h"

using namespace red_b200;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

int cuda_fail(cudaError_t err, const char *what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(err);
  return RED_B200_ECUDA;
}

#define CUDA_TRY(expr)                                   \
  do {                                                   \
    cudaError_t e_ = (expr);                             \
    if (e_ != cudaSuccess)                               \
      return cuda_fail(e_, #expr);                       \
  } while (0)

bool usable_device(int dev) {
  // asked on every call: cached (cudaGetDeviceProperties takes milliseconds)
  static std::atomic<int> known[64];  // 0 unknown, 1 usable, 2 not
  if ((dev < 0) || (dev >= 64))
    return false;
  int k = known[dev].load(std::memory_order_relaxed);
  if (k == 0) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      return false;
    k = (major == 10) ? 1 : 2;  // sm_100a code only
    known[dev].store(k, std::memory_order_relaxed);
  }
  return k == 1;
}

bool valid_style(int style) { return (style >= 1) && (style <= 5); }
bool valid_verb(int verb) { return (verb >= 0) && (verb <= 3); }

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess)
      prev = -1;
    ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (prev >= 0)
      cudaSetDevice(prev);
  }
};

} // anonymous

/* Pinned staging + device buffers for the *_host entry points: two slots so
   that the copy-in of chunk i+1 overlaps the kernel of chunk i and the
   copy-out of chunk i-1. */
struct HostPipe {
  static constexpr int kSlots = 3;
  size_t textCap = 0, lineCap = 0;
  cudaStream_t stream[kSlots]{};
  cudaEvent_t done[kSlots]{};
  uint8_t *dText[kSlots]{};
  uint64_t *dStarts[kSlots]{};
  uint32_t *dLens[kSlots]{};
  int32_t *dRes[kSlots]{};
  uint32_t *dMs[kSlots]{};
  uint32_t *dMe[kSlots]{};
  // optimistic chunks: the device word the kernels take their line count from
  // (all ones = the chunk's span held every line, 0 = it did not: nothing was
  // walked), and where it is copied back to (pinned)
  uint64_t *dN[kSlots]{};
  uint64_t *hN = nullptr;
  uint64_t slotFirst[kSlots]{}, slotCount[kSlots]{};
  bool slotOptimistic[kSlots]{};

  void release() {
    for (int s = 0; s < kSlots; ++s) {
      if (stream[s]) cudaStreamDestroy(stream[s]);
      if (done[s]) cudaEventDestroy(done[s]);
      cudaFree(dText[s]); cudaFree(dStarts[s]); cudaFree(dLens[s]);
      cudaFree(dRes[s]); cudaFree(dMs[s]); cudaFree(dMe[s]); cudaFree(dN[s]);
    }
    if (hN) cudaFreeHost(hN);
  }
};

/* red_b200_match() on one short text: a staging buffer of its own */
constexpr size_t kSingleTextMax = 256u << 10;
struct SingleMeta {
  uint64_t start;
  uint32_t len;
  int32_t result;
  uint32_t mstart, mend;
  uint32_t pad[2];  // the text behind it starts on a 16-byte boundary
};
static_assert(sizeof(SingleMeta) % 16 == 0, "text alignment");
struct SinglePipe {
  cudaStream_t stream = nullptr;
  uint8_t *hBuf = nullptr;  // pinned
  uint8_t *dBuf = nullptr;
};

/* Every line of the chunk inside [lo, hi)?  If not, *dN becomes 0 and the
   kernels that follow on the stream walk nothing. */
__global__ void span_check_kernel(const uint64_t *starts, const uint32_t *lens, uint64_t count,
                                  uint64_t lo, uint64_t hi, unsigned long long *dN) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  bool bad = false;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const uint64_t a = starts[i], b = a + lens[i];
    bad |= (a < lo) || (b > hi) || (b < a);
  }
  if (__any_sync(0xffffffffu, bad) && ((threadIdx.x & 31) == 0))
    atomicMin(dN, 0ull);
}

/* (Re)builds a hot image of the staged batch kernel from entry counts (or
   breadth-first order when there are none) and puts it on the device.  Buffers
   of an earlier image stay allocated until the program is freed: kernels
   launched with them may still be running on other streams. */
static cudaError_t upload_hot(red_b200_program *prog, const uint64_t *visits, bool asciiRows, bool acceptSticks,
                              HotDev &out, uint32_t wideStates = 0) {
  HotImage hot;
  build_hot_image(prog->parsed, prog->img, visits, wideStates ? wideStates : prog->hotMaxStates, asciiRows,
                  acceptSticks, hot, wideStates != 0);
  void *dTab = nullptr, *dC2v = nullptr, *dCodeOf = nullptr;
  cudaError_t err;
  if ((err = cudaMalloc(&dTab, hot.table.size())) != cudaSuccess)
    return err;
  prog->hotBuffers.push_back(dTab);
  if ((err = cudaMalloc(&dC2v, hot.codeToV.size() * 4)) != cudaSuccess)
    return err;
  prog->hotBuffers.push_back(dC2v);
  if ((err = cudaMalloc(&dCodeOf, hot.codeOf.size())) != cudaSuccess)
    return err;
  prog->hotBuffers.push_back(dCodeOf);
  if ((err = cudaMemcpy(dTab, hot.table.data(), hot.table.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
    return err;
  if ((err = cudaMemcpy(dC2v, hot.codeToV.data(), hot.codeToV.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
    return err;
  if ((err = cudaMemcpy(dCodeOf, hot.codeOf.data(), hot.codeOf.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
    return err;
  if ((err = cudaDeviceSynchronize()) != cudaSuccess)
    return err;
  HotDev hd;
  hd.table = static_cast<const uint8_t *>(dTab);
  hd.codeToV = static_cast<const uint32_t *>(dC2v);
  hd.codeOf = static_cast<const uint8_t *>(dCodeOf);
  hd.tableBytes = static_cast<uint32_t>(hot.table.size());
  hd.stride = hot.stride;
  hd.entryBytes = hot.entryBytes;
  hd.nHot = hot.nHot;
  hd.cAccept = hot.cAccept;
  hd.cAbs = hot.cAbs;
  hd.cPure = hot.cPure;
  hd.initCode = hot.initCode;
  hd.coverage = hot.coverage;
  hd.acceptSticks = hot.acceptSticks ? 1 : 0;
  hd.wideTable = prog->smemTable ? 0 : 1;
  out = hd;
  return cudaSuccess;
}

/* First sizeable batch of a program: walk a sample of its lines on the full
   table, count how often every state is entered, and choose the hot sets from
   that (the states a walk over THIS kind of text spends its time in): one for
   style Instant, whose walks end at the first accepting state, one for the
   styles that go on.  Synchronises the stream once.  Skipped while the stream
   is being captured. */
static void profile_hot(red_b200_program *prog, const BatchArgs &args, cudaStream_t stream,
                        bool realLines = true) {
  std::lock_guard<std::mutex> lock(prog->hotMu);
  if (prog->hotProfiled.load())
    return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if ((cudaStreamIsCapturing(stream, &cap) != cudaSuccess) || (cap != cudaStreamCaptureStatusNone)) {
    cudaGetLastError();
    return;
  }
  prog->hotProfiled.store(1);  // one attempt, whatever comes of it
  const uint32_t nStates = prog->img.stateCount;
  const size_t words = 2 * static_cast<size_t>(nStates) + 2;
  unsigned long long *dHist = nullptr;
  if (cudaMalloc(&dHist, words * 8) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  std::vector<uint64_t> visits(words, 0);
  const uint64_t stride = std::max<uint64_t>(1, args.nLines / 32768);
  bool ok = cudaMemsetAsync(dHist, 0, words * 8, stream) == cudaSuccess;
  ok = ok && (launch_profile(prog->dev, prog->smemTable ? 0 : 1, args, stride, nStates, dHist,
                             dHist + 2 * static_cast<size_t>(nStates), stream) == cudaSuccess);
  ok = ok && (cudaMemcpyAsync(visits.data(), dHist, words * 8, cudaMemcpyDeviceToHost, stream) == cudaSuccess);
  ok = ok && (cudaStreamSynchronize(stream) == cudaSuccess);
  cudaFree(dHist);
  if (!ok) {
    cudaGetLastError();
    return;
  }
  prog->launches += 1;
  uint64_t seen = 0;
  for (uint32_t st = 0; st < nStates; ++st) {
    seen += visits[st];
    visits[nStates + st] += visits[st];  // the second half becomes the whole line's count
  }
  if (seen == 0) {
    prog->hotProfiled.store(0);  // nothing was walked (a chunk whose span check failed): try again
    return;
  }
  const uint64_t samples = (args.nLines + stride - 1) / stride;
  if (samples && realLines)
    prog->avgLineBytes.store(static_cast<uint32_t>(std::min<uint64_t>(visits[2 * static_cast<size_t>(nStates)] / samples, 1u << 20)));
  {
    /* Rows of 128 columns (when nearly all of the text is ASCII) are half the
       shared memory - but every 16-byte chunk then has to be looked at for a
       byte >= 128, and that costs the batch kernel more than a sixth slot of
       text brings (measured: 256-column rows +2 %).  So the batch images take
       them only when told to (RED_B200_HOT_ASCII=1); the long-stream matcher's
       two-byte image, whose size decides how many states it holds, does when
       the text allows. */
    const uint64_t allBytes = visits[2 * static_cast<size_t>(nStates)];
    const uint64_t highBytes = visits[2 * static_cast<size_t>(nStates) + 1];
    const bool asciiText = prog->hotAsciiRows ? (prog->hotAsciiRows > 0) : (highBytes * 2000 <= allBytes);
    const bool ascii = prog->hotAsciiRows > 0;
    HotDev instant, whole;
    if ((upload_hot(prog, visits.data(), ascii, true, instant) == cudaSuccess) &&
        (upload_hot(prog, visits.data() + nStates, ascii, false, whole) == cudaSuccess)) {
      prog->hotDev[0] = instant;
      prog->hotDev[1] = whole;
      // the long-stream matcher's: every state the text enters, two-byte codes
      // if a byte cannot name them all (as many as fit beside the text stages)
      prog->hotStream = whole;
      if (whole.coverage < 1000) {
        const uint32_t rowBytes = (asciiText ? 128u : 256u) * 2 + 4;
        uint32_t fit = (prog->geo.maxSmemOptin > 140 * 1024) ? (prog->geo.maxSmemOptin - 140 * 1024) / rowBytes : 0;
        if (prog->hotStreamStates)  // RED_B200_STREAM_HOT_STATES (tests: lanes that leave the image)
          fit = std::min(fit, std::max(prog->hotStreamStates, kHotMax + 2) + 1);
        HotDev wide;
        if ((fit > kHotMax + 1) &&
            (upload_hot(prog, visits.data() + nStates, asciiText, false, wide, std::min(fit - 1, kHotMaxWide)) == cudaSuccess))
          prog->hotStream = wide;
        else
          cudaGetLastError();
      }
    }
    else
      cudaGetLastError();
  }
  /* A table in global memory: the same DFA once more with the states this text
     enters most in front, so that the rows the batch kernels mirror in shared
     memory are the ones that are used (7000 keywords, 46 613 states: the
     syslog corpus enters 746 of them; a mirror in serialized order held a
     third of the lookups). */
  if (!prog->smemTable && !prog->haveDevBatch && !getenv("RED_B200_NO_BATCH_ORDER")) {
    std::vector<uint64_t> byOld(nStates, 0);
    for (uint32_t st = 0; st < nStates; ++st)
      byOld[prog->img.newToOld[st]] = visits[nStates + st];
    DeviceImage img2;
    build_image(prog->parsed, prog->smemBudget, img2, getenv("RED_B200_NO_BYTE_ROWS") == nullptr, byOld.data());
    if ((img2.mode != kTableSmemU16) && (img2.vPerState == prog->img.vPerState)) {
      const size_t tabBytes = img2.tableBytes();
      const size_t padded = (tabBytes + 15) & ~size_t(15);
      void *dTab = nullptr, *dRes = nullptr;
      bool ok2 = (cudaMalloc(&dTab, padded) == cudaSuccess);
      if (ok2)
        prog->hotBuffers.push_back(dTab);
      ok2 = ok2 && (cudaMalloc(&dRes, img2.results.size() * 4) == cudaSuccess);
      if (ok2)
        prog->hotBuffers.push_back(dRes);
      ok2 = ok2 && (cudaMemset(dTab, 0, padded) == cudaSuccess) &&
            (cudaMemcpy(dTab, img2.table32.data(), tabBytes, cudaMemcpyHostToDevice) == cudaSuccess) &&
            (cudaMemcpy(dRes, img2.results.data(), img2.results.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess) &&
            (cudaDeviceSynchronize() == cudaSuccess);
      if (ok2) {
        ProgramDev pd = prog->dev;
        pd.table = dTab;
        pd.results = static_cast<const int32_t *>(dRes);
        pd.initV = img2.initV;
        pd.acceptV = img2.acceptV;
        pd.absV = img2.absV;
        pd.pureV = img2.pureV;
        pd.leaderV = img2.leaderV;
        prog->devBatch = pd;
        prog->haveDevBatch = true;
      }
      else
        cudaGetLastError();
    }
  }
}

extern "C" {

const char *red_b200_last_error(void) { return g_err.c_str(); }

int red_b200_abi_version(void) { return RED_B200_ABI_VERSION; }

int red_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int good = 0;
  for (int d = 0; d < n; ++d)
    if (usable_device(d))
      ++good;
  return good;
}

const char *red_b200_check_header(const void *serialized, size_t len) {
  return check_header(serialized, len);
}

int red_b200_program_upload(const void *serialized, size_t len, int device,
                            red_b200_program **out) {
  if (!out)
    return fail(RED_B200_EAPI, "null output pointer");
  *out = nullptr;
  if (!serialized || (len == 0))
    return fail(RED_B200_EAPI, "serialized dfa is empty");

  auto prog = new red_b200_program;
  std::string why;
  if (!parse_program(serialized, len, prog->parsed, why)) {
    delete prog;
    return fail(RED_B200_EAPI, why);
  }
  if (static_cast<uint64_t>(prog->parsed.stateCount) *
          prog->parsed.classCount > 0xffffffffull) {
    delete prog;
    return fail(RED_B200_ELIMIT, "dfa too big for device table");
  }

  int ndev = 0;
  if ((cudaGetDeviceCount(&ndev) != cudaSuccess) || (ndev == 0)) {
    cudaGetLastError();
    delete prog;
    return fail(RED_B200_ENODEVICE,
                "no CUDA device visible; red_b200 has no CPU fallback");
  }
  if ((device < 0) || (device >= ndev) || !usable_device(device)) {
    delete prog;
    return fail(RED_B200_ENODEVICE,
                "device is not an sm_100 GPU; red_b200 has no CPU fallback");
  }
  prog->device = device;
  DeviceGuard guard(device);

  cudaDeviceProp prop;
  cudaError_t err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) {
    delete prog;
    return cuda_fail(err, "cudaGetDeviceProperties");
  }
  prog->geo.smCount = prop.multiProcessorCount;
  prog->geo.maxSmemOptin = static_cast<int>(prop.sharedMemPerBlockOptin);

  // leave room for the class table and a little slack
  size_t budget = prop.sharedMemPerBlockOptin - 1024;
  if (const char *env = getenv("RED_B200_SMEM_TABLE_LIMIT"))
    budget = std::min<size_t>(budget, strtoull(env, nullptr, 0));
  prog->smemBudget = budget;
  build_image(prog->parsed, budget, prog->img,
              getenv("RED_B200_NO_BYTE_ROWS") == nullptr);
  DeviceImage &img = prog->img;
  prog->smemTable = (img.mode == kTableSmemU16);
  prog->format = prog->parsed.format;
  prog->leaderLen = prog->parsed.leaderLen;

  const void *tab = prog->smemTable
      ? static_cast<const void *>(img.table16.data())
      : static_cast<const void *>(img.table32.data());
  size_t tabBytes = img.tableBytes();
  size_t padded = (tabBytes + 15) & ~size_t(15);
  prog->tableBytes = tabBytes;

  auto cleanup = [&](cudaError_t e, const char *what) {
    cudaFree(prog->dTable); cudaFree(prog->dCls); cudaFree(prog->dResults);
    cudaFree(prog->dLeader);
    for (void *p : prog->hotBuffers)
      cudaFree(p);
    delete prog;
    return cuda_fail(e, what);
  };
  if ((err = cudaMalloc(&prog->dTable, padded)) != cudaSuccess)
    return cleanup(err, "cudaMalloc(table)");
  if ((err = cudaMemset(prog->dTable, 0, padded)) != cudaSuccess)
    return cleanup(err, "cudaMemset(table)");
  if ((err = cudaMemcpy(prog->dTable, tab, tabBytes, cudaMemcpyHostToDevice)) != cudaSuccess)
    return cleanup(err, "cudaMemcpy(table)");
  if ((err = cudaMalloc(&prog->dCls, 512)) != cudaSuccess)
    return cleanup(err, "cudaMalloc(cls)");
  if ((err = cudaMemcpy(prog->dCls, img.clsOffset.data(), 512, cudaMemcpyHostToDevice)) != cudaSuccess)
    return cleanup(err, "cudaMemcpy(cls)");
  if ((err = cudaMalloc(&prog->dResults, img.results.size() * 4)) != cudaSuccess)
    return cleanup(err, "cudaMalloc(results)");
  if ((err = cudaMemcpy(prog->dResults, img.results.data(), img.results.size() * 4,
                        cudaMemcpyHostToDevice)) != cudaSuccess)
    return cleanup(err, "cudaMemcpy(results)");

  {
    // raw class map (256 bytes) followed by the leader (classes)
    std::vector<uint8_t> aux(256 + std::max<size_t>(1, img.leader.size()), 0);
    memcpy(aux.data(), prog->parsed.equiv, 256);
    for (size_t ii = 0; ii < img.leader.size(); ++ii)
      aux[256 + ii] = img.leader[ii];
    if ((err = cudaMalloc(&prog->dLeader, aux.size())) != cudaSuccess)
      return cleanup(err, "cudaMalloc(leader)");
    if ((err = cudaMemcpy(prog->dLeader, aux.data(), aux.size(),
                          cudaMemcpyHostToDevice)) != cudaSuccess)
      return cleanup(err, "cudaMemcpy(leader)");
  }

  {
    // [256 * 8 reset states | 256 reset counts]
    const size_t statesBytes = img.resetStates.size() * 4;
    void *dReset = nullptr;
    if ((err = cudaMalloc(&dReset, statesBytes + 256)) != cudaSuccess)
      return cleanup(err, "cudaMalloc(reset states)");
    prog->hotBuffers.push_back(dReset);
    if ((err = cudaMemcpy(dReset, img.resetStates.data(), statesBytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (err = cudaMemcpy(static_cast<char *>(dReset) + statesBytes, img.resetCount.data(), 256,
                          cudaMemcpyHostToDevice)) != cudaSuccess) {
      return cleanup(err, "cudaMemcpy(reset states)");
    }
    prog->dev.resetStates = static_cast<const uint32_t *>(dReset);
    prog->dev.resetCount = reinterpret_cast<const uint8_t *>(dReset) + statesBytes;
  }

  {
    // the start-position filter of the sliding verbs (device_program.h)
    const ParsedProgram &pp = prog->parsed;
    const uint32_t C = pp.classCount;
    std::vector<uint32_t> bits(2048 + 8, 0);
    for (uint32_t b0 = 0; b0 < 256; ++b0) {
      const uint32_t v1 = pp.next[static_cast<size_t>(pp.initialState) * C + pp.equiv[b0]];
      const bool acc1 = !pp.pureDead[v1] && (pp.result[v1] > 0);
      if (acc1)
        bits[2048 + (b0 >> 5)] |= 1u << (b0 & 31);
      for (uint32_t b1 = 0; b1 < 256; ++b1) {
        const uint32_t v2 = pp.next[static_cast<size_t>(v1) * C + pp.equiv[b1]];
        if (acc1 || (!pp.pureDead[v1] && !pp.pureDead[v2])) {
          const uint32_t idx = (b0 << 8) | b1;
          bits[idx >> 5] |= 1u << (idx & 31);
        }
      }
    }
    void *dFilter = nullptr;
    if ((err = cudaMalloc(&dFilter, bits.size() * 4)) != cudaSuccess)
      return cleanup(err, "cudaMalloc(start filter)");
    prog->hotBuffers.push_back(dFilter);
    if ((err = cudaMemcpy(dFilter, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
      return cleanup(err, "cudaMemcpy(start filter)");
    }
    prog->dev.startFilter = static_cast<const uint32_t *>(dFilter);
  }

  ProgramDev &pd = prog->dev;
  pd.equiv = static_cast<const uint8_t *>(prog->dLeader);
  pd.leader = pd.equiv + 256;
  pd.leaderLen = static_cast<uint32_t>(img.leader.size());
  pd.leaderV = img.leaderV;
  pd.clsOffset = static_cast<const uint16_t *>(prog->dCls);
  pd.table = prog->dTable;
  pd.results = static_cast<const int32_t *>(prog->dResults);
  pd.tableBytes = static_cast<uint32_t>(padded);
  pd.unitShift = img.unitShift;
  pd.initV = img.initV;
  pd.acceptV = img.acceptV;
  pd.absV = img.absV;
  pd.pureV = img.pureV;
  pd.vPerState = img.vPerState;
  pd.stateCount = img.stateCount;
  {
    // reciprocal for v / vPerState, kept only if exact over the whole v range
    const uint64_t d = img.vPerState;
    const uint64_t maxV = static_cast<uint64_t>(img.stateCount) * d;
    uint64_t magic = (1ull << 32) / d + 1;
    bool exact = (magic <= 0xffffffffull) && (maxV <= (1ull << 20));
    for (uint64_t v = 0; exact && (v < maxV); ++v)
      exact = ((v * magic) >> 32) == v / d;
    pd.divMagic = exact ? static_cast<uint32_t>(magic) : 0;
  }
  pd.clsMaxOffset = *std::max_element(img.clsOffset.begin(), img.clsOffset.end());
  pd.initResult = img.initResult;
  pd.byteRows = img.byteRows ? 1 : 0;
  pd.hasFinalStates = img.finalReachable ? 1 : 0;
  pd.hotBytes = 0;
  if (!prog->smemTable) {
    // rows of the shallowest states (the renumbering keeps the serialized,
    // breadth-first order inside each group) go to shared memory as well.
    // Not all of it: what shared memory takes, L1 loses, and the rest of the
    // table and the text live there.  Measured on B200 (2000 keywords, 2.2 MB
    // table, tools/hot_rows_sweep.py): no mirror 886 GB/s, 16 KB 1017,
    // 96 KB 1046, 160 KB 1070, 203 KB 847.
    size_t room = prop.sharedMemPerBlockOptin - (68u << 10);
    if (const char *env = getenv("RED_B200_HOT_ROWS_BYTES"))
      room = std::min<size_t>(room, strtoull(env, nullptr, 0));
    pd.hotBytes = static_cast<uint32_t>(std::min<size_t>(room, padded) & ~size_t(15));
  }

  // cudaMemcpy from pageable memory may return before the DMA has landed, and
  // the kernels run on non-blocking streams that do not order against the
  // legacy default stream: make the upload visible before anyone can launch.
  if ((err = cudaDeviceSynchronize()) != cudaSuccess)
    return cleanup(err, "cudaDeviceSynchronize(upload)");

  // the host copies of the big tables are no longer needed
  std::vector<uint16_t>().swap(img.table16);
  std::vector<uint32_t>().swap(img.table32);

  if (const char *env = getenv("RED_B200_THREADS"))
    prog->tune.threads = atoi(env);
  if (const char *env = getenv("RED_B200_CTAS_PER_SM"))
    prog->tune.ctasPerSm = atoi(env);
  if (const char *env = getenv("RED_B200_VARIANT"))
    prog->tune.variant = atoi(env);
  if (const char *env = getenv("RED_B200_LINES_PER_LANE"))
    prog->tune.linesPerLane = atoi(env);
  if (const char *env = getenv("RED_B200_SIMPLE_WALK"))
    prog->tune.simpleWalk = atoi(env) != 0;

  // the staged batch kernel's image: breadth-first hot set until a batch has
  // been profiled
  if (const char *env = getenv("RED_B200_HOT"))
    prog->hotEnabled = atoi(env) != 0;
  if (const char *env = getenv("RED_B200_HOT_STATES"))
    prog->hotMaxStates = std::max(1, std::min(atoi(env), static_cast<int>(kHotMax)));
  if (const char *env = getenv("RED_B200_HOT_PROFILE"))
    prog->hotProfileWanted = atoi(env) != 0;
  if (const char *env = getenv("RED_B200_HOT_WARPS"))
    prog->hotTune.consumerWarps = atoi(env);
  if (const char *env = getenv("RED_B200_HOT_DEBUG"))
    if (atoi(env) != 0) {
      void *dbg = nullptr;
      if (cudaMalloc(&dbg, 16 * 8) == cudaSuccess) {
        cudaMemset(dbg, 0, 16 * 8);
        prog->hotBuffers.push_back(dbg);
        prog->hotTune.debug = static_cast<unsigned long long *>(dbg);
      }
    }
  if (const char *env = getenv("RED_B200_HOT_ASCII"))
    prog->hotAsciiRows = atoi(env);
  if (const char *env = getenv("RED_B200_HOT_FORCE"))
    prog->hotTune.force = atoi(env) != 0;
  if (const char *env = getenv("RED_B200_HOT_DEFER"))
    prog->hotTune.defer = atoi(env);
  if (const char *env = getenv("RED_B200_HOT_PRODUCERS"))
    prog->hotTune.producers = atoi(env);
  if (const char *env = getenv("RED_B200_HOT_SLOTS"))
    prog->hotTune.slots = atoi(env);
  if (const char *env = getenv("RED_B200_HOT_BATCH"))
    prog->hotTune.batchLines = atoi(env);
  if (const char *env = getenv("RED_B200_HOT_SLOT_BYTES"))
    prog->hotTune.slotBytes = atoi(env);
  if (const char *env = getenv("RED_B200_STREAM_HOT_STATES"))
    prog->hotStreamStates = static_cast<uint32_t>(std::max(0, atoi(env)));
  if (prog->hotEnabled) {
    if (((err = upload_hot(prog, nullptr, prog->hotAsciiRows > 0, true, prog->hotDev[0])) != cudaSuccess) ||
        ((err = upload_hot(prog, nullptr, prog->hotAsciiRows > 0, false, prog->hotDev[1])) != cudaSuccess))
      return cleanup(err, "upload of the hot image");
    prog->hotStream = prog->hotDev[1];
  }

  *out = prog;
  return RED_B200_OK;
}

void red_b200_program_free(red_b200_program *prog) {
  if (!prog)
    return;
  DeviceGuard guard(prog->device);
  if (prog->pipe) {
    prog->pipe->release();
    delete prog->pipe;
  }
  release_grep_pipe(prog);
  if (prog->single) {
    if (prog->single->stream) cudaStreamDestroy(prog->single->stream);
    if (prog->single->hBuf) cudaFreeHost(prog->single->hBuf);
    cudaFree(prog->single->dBuf);
    delete prog->single;
  }
  for (void *p : prog->hotBuffers)
    cudaFree(p);
  cudaFree(prog->dTable);
  cudaFree(prog->dCls);
  cudaFree(prog->dResults);
  cudaFree(prog->dLeader);
  delete prog;
}

int red_b200_program_get_info(const red_b200_program *prog,
                              red_b200_program_info *info) {
  if (!prog || !info)
    return fail(RED_B200_EAPI, "null argument");
  memset(info, 0, sizeof(*info));
  info->abi_version = RED_B200_ABI_VERSION;
  info->format = prog->format;
  info->state_count = prog->img.stateCount;
  info->class_count = prog->img.classCount;
  info->leader_len = prog->leaderLen;
  info->table_mode = prog->smemTable ? RED_B200_TABLE_SMEM_U16
                                     : RED_B200_TABLE_GLOBAL_U32;
  info->row_stride = prog->img.rowStride;
  info->accepting_states = prog->img.acceptCount;
  info->pure_dead_states = prog->img.pureCount;
  info->max_result = static_cast<uint32_t>(prog->img.maxResult);
  info->table_bytes = prog->tableBytes;
  info->device = prog->device;
  info->sm_count = prog->geo.smCount;
  return RED_B200_OK;
}

/* kernels launched through this program so far (bench.py's gpu_launches) */
long red_b200_program_launch_count(const red_b200_program *prog) {
  return prog ? prog->launches.load() : 0;
}

/* RED_B200_HOT_DEBUG=1 at upload: cycle counters of the staged batch kernel,
   summed over all launches so far (synchronises the device; zeroes them).
   [0..3] producers: total, waiting for index entries, waiting for a slot, batches;
   [4..8] consumers: total, waiting for a batch, walking, writing back, tiles. */
int red_b200_program_debug_counters(red_b200_program *prog, uint64_t out[16]) {
  if (!prog || !out)
    return fail(RED_B200_EAPI, "null argument");
  memset(out, 0, 16 * 8);
  if (!prog->hotTune.debug)
    return RED_B200_OK;
  DeviceGuard guard(prog->device);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out, prog->hotTune.debug, 16 * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemset(prog->hotTune.debug, 0, 16 * 8));
  return RED_B200_OK;
}

/* of those, launches of the staged batch kernel (kernels_hot.cu) */
long red_b200_program_hot_launch_count(const red_b200_program *prog) {
  return prog ? prog->hotLaunches.load() : 0;
}

static int launch_batch(red_b200_program *prog, const BatchArgs &args,
                        int verb, int style, cudaStream_t stream) {
  int launches = 0;
  cudaError_t err;
  /* Every launch gets a work counter of its own, from the stream-ordered pool:
     it is handed back behind the kernel on the same stream, so it cannot be
     zeroed for another launch while this one still counts on it (launches of
     one program overlap on different streams). */
  unsigned long long *ctr = nullptr;
  if ((err = scratch_alloc(&ctr, sizeof(unsigned long long), stream)) != cudaSuccess)
    return cuda_fail(err, "work counter");
  struct CounterRelease {
    unsigned long long *p;
    cudaStream_t st;
    ~CounterRelease() { scratch_free(p, st); }
  } counterRelease{ctr, stream};
  if ((verb == RED_B200_VERB_CHECK) || (verb == RED_B200_VERB_MATCH)) {
    BatchArgs a = args;
    a.checkLeader = (verb == RED_B200_VERB_CHECK) && (prog->dev.leaderLen > 0);
    // the staged kernel (kernels_hot.cu) where it applies
    err = cudaErrorNotSupported;
    if (prog->hotEnabled && (prog->tune.variant == 0) && !a.checkLeader && !a.matchStarts &&
        ((style == RED_B200_STY_INSTANT) || (style == RED_B200_STY_LAST) || (style == RED_B200_STY_FULL))) {
      if (prog->hotProfileWanted && !prog->hotProfiled.load() && (a.nLines >= 8192))
        profile_hot(prog, a, stream);
      HotDev hd;
      {
        std::lock_guard<std::mutex> lock(prog->hotMu);
        hd = prog->hotDev[(style == RED_B200_STY_INSTANT) ? 0 : 1];
      }
      err = launch_batch_hot(prog->dev, hd, a, style, prog->geo, prog->hotTune,
                             prog->avgLineBytes.load(), ctr, stream, &launches);
      if (err == cudaSuccess)
        prog->hotLaunches += 1;
    }
    if (err != cudaErrorNotSupported)
      ;
    else {
      // a table in global memory: the image ordered by what the text enters
      // (made by the first sizeable batch's profile, whatever its style)
      if (!prog->smemTable && prog->hotProfileWanted && !prog->hotProfiled.load() && (a.nLines >= 8192))
        profile_hot(prog, a, stream);
      ProgramDev pdUse = prog->dev;
      if (!prog->smemTable) {
        std::lock_guard<std::mutex> lock(prog->hotMu);
        if (prog->haveDevBatch)
          pdUse = prog->devBatch;
      }
      if (prog->tune.variant == 1)  // v1 tile kernel, kept as the A/B baseline
        err = launch_batch_anchored(pdUse, prog->smemTable, a, style,
                                    prog->geo, prog->tune, stream, &launches);
      else
        err = launch_batch_lanes(pdUse, prog->smemTable, a, style, prog->geo,
                                 prog->tune, ctr, stream, &launches);
    }
  }
  else {
    err = launch_batch_sliding(prog->dev, prog->smemTable, args, style,
                               verb == RED_B200_VERB_SEARCH, prog->geo, stream,
                               &launches, ctr, prog->tune.simpleWalk);
  }
  if (err != cudaSuccess)
    return cuda_fail(err, "batch kernel launch");
  prog->launches += launches;
  return RED_B200_OK;
}

} // extern "C"

namespace red_b200 {
void set_last_error(const std::string &msg) { g_err = msg; }

int launch_batch_for(red_b200_program *prog, const BatchArgs &args, int verb,
                     int style, cudaStream_t stream, std::string &why) {
  int rc = launch_batch(prog, args, verb, style, stream);
  if (rc != RED_B200_OK)
    why = g_err;
  return rc;
}
}

extern "C" {

int red_b200_match_batch_device(const red_b200_program *cprog,
                                const uint8_t *d_text,
                                const uint64_t *d_starts,
                                const uint32_t *d_lens, uint64_t n_lines,
                                int verb, int style, int32_t *d_results,
                                uint32_t *d_match_starts,
                                uint32_t *d_match_ends, void *stream) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog)
    return fail(RED_B200_EAPI, "null program");
  if (!valid_style(style))
    return fail(RED_B200_EEXEC, "unsupported style");
  if (!valid_verb(verb))
    return fail(RED_B200_EEXEC, "unsupported verb");
  if (n_lines == 0)
    return RED_B200_OK;
  if (!d_text || !d_starts || !d_lens || !d_results)
    return fail(RED_B200_EAPI, "null buffer");
  BatchArgs args;
  args.text = d_text;
  args.starts = d_starts;
  args.lens = d_lens;
  args.nLines = n_lines;
  args.results = d_results;
  // check/scan report no positions (Matcher.h:74-77, 84-87)
  bool positions = (verb == RED_B200_VERB_MATCH) || (verb == RED_B200_VERB_SEARCH);
  args.matchStarts = positions ? d_match_starts : nullptr;
  args.matchEnds = positions ? d_match_ends : nullptr;
  DeviceGuard guard(prog->device);
  return launch_batch(prog, args, verb, style, static_cast<cudaStream_t>(stream));
}

static int ensure_pipe(red_b200_program *prog, size_t textCap, size_t lineCap) {
  HostPipe *pp = prog->pipe;
  if (pp && (pp->textCap >= textCap) && (pp->lineCap >= lineCap))
    return RED_B200_OK;
  if (pp) {
    pp->release();
    delete pp;
    prog->pipe = nullptr;
  }
  pp = new HostPipe;
  pp->textCap = textCap;
  pp->lineCap = lineCap;
  prog->pipe = pp;
  for (int s = 0; s < HostPipe::kSlots; ++s) {
    CUDA_TRY(cudaStreamCreateWithFlags(&pp->stream[s], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&pp->done[s], cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&pp->dText[s], textCap + 64));
    CUDA_TRY(cudaMalloc(&pp->dStarts[s], lineCap * 8));
    CUDA_TRY(cudaMalloc(&pp->dLens[s], lineCap * 4));
    CUDA_TRY(cudaMalloc(&pp->dRes[s], lineCap * 4));
    CUDA_TRY(cudaMalloc(&pp->dMs[s], lineCap * 4));
    CUDA_TRY(cudaMalloc(&pp->dMe[s], lineCap * 4));
    CUDA_TRY(cudaMalloc(&pp->dN[s], 8));
  }
  CUDA_TRY(cudaHostAlloc(&pp->hN, HostPipe::kSlots * 8, cudaHostAllocDefault));
  return RED_B200_OK;
}

int red_b200_match_batch_host(const red_b200_program *cprog,
                              const uint8_t *text, uint64_t text_len,
                              const uint64_t *starts, const uint32_t *lens,
                              uint64_t n_lines, int verb, int style,
                              int32_t *results, uint32_t *match_starts,
                              uint32_t *match_ends) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog)
    return fail(RED_B200_EAPI, "null program");
  if (!valid_style(style))
    return fail(RED_B200_EEXEC, "unsupported style");
  if (!valid_verb(verb))
    return fail(RED_B200_EEXEC, "unsupported verb");
  if (n_lines == 0)
    return RED_B200_OK;
  if (!text || !starts || !lens || !results)
    return fail(RED_B200_EAPI, "null buffer");
  bool positions = (verb == RED_B200_VERB_MATCH) || (verb == RED_B200_VERB_SEARCH);
  if (!positions) {
    match_starts = nullptr;
    match_ends = nullptr;
  }

  std::lock_guard<std::mutex> lock(prog->mu);
  DeviceGuard guard(prog->device);

  /* Cut the batch into chunks of consecutive lines whose bytes form one
     span of the text (lines are normally in text order; a batch that is not
     is still correct, its spans just get longer).  Chunk targets: ~32 MiB of
     text or 1 Mi lines, whichever comes first. */
  size_t textTarget = 32u << 20;
  size_t lineTarget = 1u << 20;
  if (const char *env = getenv("RED_B200_HOST_CHUNK_BYTES"))
    textTarget = std::max<size_t>(1 << 16, strtoull(env, nullptr, 0));

  /* Planning a chunk exactly is a min/max over its lines' extents - 12 bytes
     of index per line through the host's caches, which with several ranks on
     one host starves the copies.  Lines are normally in text order, so a
     chunk is planned OPTIMISTICALLY in O(log n): the span runs from the first
     line's start to the last line's end, the count is found by bisection on
     the starts, and a kernel checks on the device, in front of the matcher,
     that every line really lies inside the span.  If not, nothing of that
     chunk is walked (the kernels take their line count from a device word the
     check zeroes) and the chunk is redone with the exact plan afterwards. */
  const uint64_t kBlock = 4096;
  const bool optimistic = getenv("RED_B200_EXACT_PLAN") == nullptr;
  std::vector<std::pair<uint64_t, uint64_t>> redo;  // (first, count) of chunks that failed the check
  auto endOf = [&](uint64_t i) { return starts[i] + lens[i]; };

  // runs lines [from, from + todo): optimistic chunks if allowed, else exact ones
  auto run = [&](uint64_t from, uint64_t todo, bool allowOptimistic) -> int {
    uint64_t first = from;
    const uint64_t stop = from + todo;
    int rc = RED_B200_OK;
    for (size_t ci = 0; first < stop; ++ci) {
      uint64_t lo = ~0ull, hi = 0, count = 0;
      bool opt = false;
      if (allowOptimistic) {
        lo = starts[first];
        uint64_t c = std::min<uint64_t>(lineTarget, stop - first);
        if (endOf(first + c - 1) - lo > textTarget) {
          // largest count whose last line still ends inside the target (if the
          // starts ascend; if they do not, the device check will say so)
          uint64_t good = 1, bad = c;
          while (good + 1 < bad) {
            const uint64_t mid = good + (bad - good) / 2;
            if ((endOf(first + mid - 1) >= lo) && (endOf(first + mid - 1) - lo <= textTarget))
              good = mid;
            else
              bad = mid;
          }
          c = good;
        }
        hi = endOf(first + c - 1);
        count = c;
        opt = (hi >= lo) && (hi <= text_len) && (hi - lo <= textTarget + (8u << 20));
      }
      if (!opt) {
        lo = ~0ull;
        hi = 0;
        count = 0;
        while (first + count < stop) {
          const uint64_t m = std::min<uint64_t>(kBlock, stop - first - count);
          const uint64_t *sp = starts + first + count;
          const uint32_t *lp = lens + first + count;
          uint64_t blo = ~0ull, bhi = 0;
          for (uint64_t k = 0; k < m; ++k) {
            const uint64_t a0 = sp[k], a1 = sp[k] + lp[k];
            blo = a0 < blo ? a0 : blo;
            bhi = a1 > bhi ? a1 : bhi;
          }
          const uint64_t nlo = std::min(lo, blo), nhi = std::max(hi, bhi);
          if (count && (((nhi - nlo) > textTarget) || (count + m > lineTarget)))
            break;
          lo = nlo;
          hi = nhi;
          count += m;
        }
        if (hi > text_len)
          return fail(RED_B200_EAPI, "line extends past end of text");
      }
      // (re)size the staging buffers; growing waits for the work in flight
      if (!prog->pipe || (prog->pipe->textCap < hi - lo) || (prog->pipe->lineCap < count)) {
        if (prog->pipe) {
          for (int s = 0; s < HostPipe::kSlots; ++s)
            CUDA_TRY(cudaStreamSynchronize(prog->pipe->stream[s]));
          for (int s = 0; s < HostPipe::kSlots; ++s)
            if (prog->pipe->slotOptimistic[s] && (prog->pipe->hN[s] == 0)) {
              redo.emplace_back(prog->pipe->slotFirst[s], prog->pipe->slotCount[s]);
              prog->pipe->slotOptimistic[s] = false;
            }
        }
        rc = ensure_pipe(prog, std::max<size_t>(hi - lo, textTarget + (8u << 20)),
                         std::max<size_t>(count, lineTarget));
        if (rc != RED_B200_OK)
          return rc;
      }
      HostPipe *pp = prog->pipe;
      const int s = static_cast<int>(ci % HostPipe::kSlots);
      cudaStream_t st = pp->stream[s];
      // slot free? (its previous copy-out finished)
      CUDA_TRY(cudaEventSynchronize(pp->done[s]));
      if (pp->slotOptimistic[s] && (pp->hN[s] == 0))
        redo.emplace_back(pp->slotFirst[s], pp->slotCount[s]);
      pp->slotOptimistic[s] = opt;
      pp->slotFirst[s] = first;
      pp->slotCount[s] = count;
      pp->hN[s] = ~0ull;
      // The device copy keeps the span's alignment modulo 16, and the kernels are
      // handed a text base shifted by the span's (aligned) start, so the caller's
      // absolute start offsets are used as they are - no per-line rewriting.
      const uint64_t lo16 = lo & ~uint64_t(15);
      CUDA_TRY(cudaMemcpyAsync(pp->dStarts[s], starts + first, count * 8,
                               cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(pp->dLens[s], lens + first, count * 4,
                               cudaMemcpyHostToDevice, st));
      if (opt) {
        CUDA_TRY(cudaMemsetAsync(pp->dN[s], 0xff, 8, st));
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>((count + 255) / 256, 4 * 148));
        span_check_kernel<<<grid, 256, 0, st>>>(pp->dStarts[s], pp->dLens[s], count, lo, hi,
                                                reinterpret_cast<unsigned long long *>(pp->dN[s]));
      }
      CUDA_TRY(cudaMemcpyAsync(pp->dText[s] + (lo - lo16), text + lo, hi - lo,
                               cudaMemcpyHostToDevice, st));
      BatchArgs args;
      args.text = pp->dText[s] - lo16;
      args.starts = pp->dStarts[s];
      args.lens = pp->dLens[s];
      args.nLines = count;
      args.nLinesDev = opt ? pp->dN[s] : nullptr;
      args.results = pp->dRes[s];
      args.matchStarts = match_starts ? pp->dMs[s] : nullptr;
      args.matchEnds = match_ends ? pp->dMe[s] : nullptr;
      rc = launch_batch(prog, args, verb, style, st);
      if (rc != RED_B200_OK)
        return rc;
      CUDA_TRY(cudaMemcpyAsync(results + first, pp->dRes[s], count * 4,
                               cudaMemcpyDeviceToHost, st));
      if (match_starts)
        CUDA_TRY(cudaMemcpyAsync(match_starts + first, pp->dMs[s], count * 4,
                                 cudaMemcpyDeviceToHost, st));
      if (match_ends)
        CUDA_TRY(cudaMemcpyAsync(match_ends + first, pp->dMe[s], count * 4,
                                 cudaMemcpyDeviceToHost, st));
      if (opt)
        CUDA_TRY(cudaMemcpyAsync(pp->hN + s, pp->dN[s], 8, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaEventRecord(pp->done[s], st));
      first += count;
    }
    // drain: everything copied out, failed optimistic chunks collected
    HostPipe *pp = prog->pipe;
    if (pp) {
      for (int s = 0; s < HostPipe::kSlots; ++s)
        CUDA_TRY(cudaStreamSynchronize(pp->stream[s]));
      for (int s = 0; s < HostPipe::kSlots; ++s) {
        if (pp->slotOptimistic[s] && (pp->hN[s] == 0))
          redo.emplace_back(pp->slotFirst[s], pp->slotCount[s]);
        pp->slotOptimistic[s] = false;
      }
    }
    return RED_B200_OK;
  };

  // an error leaves nothing in flight: copies into the caller's buffers may
  // still be queued on the slots' streams when run() gives up
  auto settle = [&](int code) {
    if (prog->pipe) {
      const std::string keep = g_err;
      for (int s = 0; s < HostPipe::kSlots; ++s) {
        if (prog->pipe->stream[s])
          cudaStreamSynchronize(prog->pipe->stream[s]);
        prog->pipe->slotOptimistic[s] = false;
      }
      cudaGetLastError();
      g_err = keep;
    }
    return code;
  };
  int rc = run(0, n_lines, optimistic);
  if (rc != RED_B200_OK)
    return settle(rc);
  // chunks whose lines were not all inside the optimistic span: exact plan
  std::vector<std::pair<uint64_t, uint64_t>> again;
  again.swap(redo);
  for (auto &fc : again) {
    rc = run(fc.first, fc.second, false);
    if (rc != RED_B200_OK)
      return settle(rc);
  }
  return RED_B200_OK;
}

int red_b200_match(const red_b200_program *prog, const void *text, size_t len,
                   int verb, int style, red_b200_outcome *out) {
  if (!prog || !out)
    return fail(RED_B200_EAPI, "null argument");
  if (!valid_style(style))
    return fail(RED_B200_EEXEC, "unsupported style");
  if (!valid_verb(verb))
    return fail(RED_B200_EEXEC, "unsupported verb");
  if (len && !text)
    return fail(RED_B200_EAPI, "null text");
  // Long texts: copy to the GPU once and take the chunked stream path.  (check()
  // on a program with a leader keeps the per-line kernel: its leader quirk is
  // not a property of the stream matcher.)
  const bool streamable = (verb == RED_B200_VERB_MATCH) ||
                          ((verb == RED_B200_VERB_CHECK) && (prog->leaderLen == 0));
  if (streamable && (len >= (256u << 10))) {
    DeviceGuard guard(prog->device);
    uint8_t *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, len + 64));
    cudaError_t e = cudaMemcpy(d, text, len, cudaMemcpyHostToDevice);
    int rc = (e == cudaSuccess)
        ? red_b200_match_stream_device(prog, d, len, verb, style, out, nullptr)
        : cuda_fail(e, "cudaMemcpy(text)");
    cudaFree(d);
    return rc;
  }
  if (len >= (1ull << 31))
    return fail(RED_B200_ELIMIT, "text too long for this verb; use match on the stream entry point");
  if (len <= kSingleTextMax) {
    // a short text: one pinned staging buffer, one stream, three small copies
    // and the batch kernel for a single line - not the three-slot pipeline of
    // the batch entry point (tens of megabytes of device buffers)
    red_b200_program *p = const_cast<red_b200_program *>(prog);
    std::lock_guard<std::mutex> lock(p->mu);
    DeviceGuard guard(p->device);
    if (!p->single) {
      SinglePipe *sp = new SinglePipe;
      p->single = sp;
      CUDA_TRY(cudaStreamCreateWithFlags(&sp->stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaHostAlloc(&sp->hBuf, kSingleTextMax + 64 + sizeof(SingleMeta), cudaHostAllocDefault));
      CUDA_TRY(cudaMalloc(&sp->dBuf, kSingleTextMax + 64 + sizeof(SingleMeta)));
    }
    SinglePipe *sp = p->single;
    // [meta | text]: one copy in, the meta block back
    SingleMeta *hm = reinterpret_cast<SingleMeta *>(sp->hBuf);
    SingleMeta *dm = reinterpret_cast<SingleMeta *>(sp->dBuf);
    hm->start = 0;
    hm->len = static_cast<uint32_t>(len);
    hm->result = 0;
    hm->mstart = hm->mend = 0;
    if (len)
      memcpy(sp->hBuf + sizeof(SingleMeta), text, len);
    CUDA_TRY(cudaMemcpyAsync(sp->dBuf, sp->hBuf, sizeof(SingleMeta) + len, cudaMemcpyHostToDevice, sp->stream));
    BatchArgs args;
    args.text = sp->dBuf + sizeof(SingleMeta);
    args.starts = &dm->start;
    args.lens = &dm->len;
    args.nLines = 1;
    args.results = &dm->result;
    const bool positions = (verb == RED_B200_VERB_MATCH) || (verb == RED_B200_VERB_SEARCH);
    args.matchStarts = positions ? &dm->mstart : nullptr;
    args.matchEnds = positions ? &dm->mend : nullptr;
    int rc = launch_batch(p, args, verb, style, sp->stream);
    if (rc != RED_B200_OK)
      return rc;
    CUDA_TRY(cudaMemcpyAsync(hm, dm, sizeof(SingleMeta), cudaMemcpyDeviceToHost, sp->stream));
    CUDA_TRY(cudaStreamSynchronize(sp->stream));
    out->result = hm->result;
    out->start = hm->mstart;
    out->end = hm->mend;
    return RED_B200_OK;
  }
  uint64_t start0 = 0;
  uint32_t len32 = static_cast<uint32_t>(len);
  int32_t res = 0;
  uint32_t ms = 0, me = 0;
  static const uint8_t empty[1] = {0};
  const uint8_t *t = len ? static_cast<const uint8_t *>(text) : empty;
  int rc = red_b200_match_batch_host(prog, t, len, &start0, &len32, 1, verb,
                                     style, &res, &ms, &me);
  if (rc != RED_B200_OK)
    return rc;
  out->result = res;
  out->start = ms;
  out->end = me;
  return RED_B200_OK;
}

int red_b200_multi_batch_device(const red_b200_program *cprog,
                                const uint8_t *d_text, const uint64_t *d_starts,
                                const uint32_t *d_lens, uint64_t n_lines,
                                int which, uint32_t max_per_line,
                                uint32_t *d_counts, red_b200_span *d_spans,
                                void *stream) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog)
    return fail(RED_B200_EAPI, "null program");
  if ((which != RED_B200_MULTI_MATCH_ALL) && (which != RED_B200_MULTI_COLLECT))
    return fail(RED_B200_EEXEC, "unsupported multi-match kind");
  if (n_lines == 0)
    return RED_B200_OK;
  if (!d_text || !d_starts || !d_lens || !d_counts || (max_per_line && !d_spans))
    return fail(RED_B200_EAPI, "null buffer");
  DeviceGuard guard(prog->device);
  int launches = 0;
  static_assert(sizeof(red_b200_span) == sizeof(MultiSpan), "span layout");
  cudaError_t err = launch_batch_multi(prog->dev, prog->smemTable, d_text, d_starts, d_lens,
                                       n_lines, which, max_per_line, d_counts,
                                       reinterpret_cast<MultiSpan *>(d_spans), prog->geo,
                                       static_cast<cudaStream_t>(stream), &launches);
  if (err != cudaSuccess)
    return cuda_fail(err, "multi-match kernel launch");
  prog->launches += launches;
  return RED_B200_OK;
}

int red_b200_multi_batch_host(const red_b200_program *cprog,
                              const uint8_t *text, uint64_t text_len,
                              const uint64_t *starts, const uint32_t *lens,
                              uint64_t n_lines, int which, uint32_t max_per_line,
                              uint32_t *counts, red_b200_span *spans) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog)
    return fail(RED_B200_EAPI, "null program");
  if (n_lines == 0)
    return RED_B200_OK;
  if (!text || !starts || !lens || !counts || (max_per_line && !spans))
    return fail(RED_B200_EAPI, "null buffer");
  for (uint64_t i = 0; i < n_lines; ++i)
    if (starts[i] + lens[i] > text_len)
      return fail(RED_B200_EAPI, "line extends past end of text");
  DeviceGuard guard(prog->device);
  // simple staging: this entry point is for moderate batches (articles, lines)
  uint8_t *dText = nullptr;
  uint64_t *dStarts = nullptr;
  uint32_t *dLens = nullptr, *dCounts = nullptr;
  red_b200_span *dSpans = nullptr;
  auto freeAll = [&]() {
    cudaFree(dText); cudaFree(dStarts); cudaFree(dLens); cudaFree(dCounts); cudaFree(dSpans);
  };
  const size_t spanBytes = std::max<size_t>(1, n_lines * max_per_line) * sizeof(red_b200_span);
  cudaError_t e;
  if ((e = cudaMalloc(&dText, text_len + 64)) != cudaSuccess ||
      (e = cudaMalloc(&dStarts, n_lines * 8)) != cudaSuccess ||
      (e = cudaMalloc(&dLens, n_lines * 4)) != cudaSuccess ||
      (e = cudaMalloc(&dCounts, n_lines * 4)) != cudaSuccess ||
      (e = cudaMalloc(&dSpans, spanBytes)) != cudaSuccess ||
      (e = cudaMemcpy(dText, text, text_len, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(dStarts, starts, n_lines * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(dLens, lens, n_lines * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
    freeAll();
    return cuda_fail(e, "multi-match staging");
  }
  int rc = red_b200_multi_batch_device(prog, dText, dStarts, dLens, n_lines, which,
                                       max_per_line, dCounts, dSpans, nullptr);
  if (rc == RED_B200_OK) {
    if ((e = cudaMemcpy(counts, dCounts, n_lines * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
        (max_per_line &&
         (e = cudaMemcpy(spans, dSpans, n_lines * max_per_line * sizeof(red_b200_span),
                         cudaMemcpyDeviceToHost)) != cudaSuccess))
      rc = cuda_fail(e, "multi-match copy-out");
  }
  freeAll();
  return rc;
}

int red_b200_replace_batch_device(const red_b200_program *cprog,
                                  const uint8_t *d_text, const uint64_t *d_starts,
                                  const uint32_t *d_lens, uint64_t n_lines,
                                  const uint8_t *repl, uint32_t repl_len,
                                  uint64_t max_per_line, int style, int out_delim,
                                  uint8_t *d_out_text, uint64_t out_capacity,
                                  uint64_t *d_out_starts, uint32_t *d_out_lens,
                                  uint32_t *d_counts, uint64_t *out_bytes,
                                  void *stream) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog || !out_bytes)
    return fail(RED_B200_EAPI, "null argument");
  *out_bytes = 0;
  if (!valid_style(style))
    return fail(RED_B200_EEXEC, "unsupported style");
  if (out_delim > 255)
    return fail(RED_B200_EAPI, "delimiter is not a byte");
  if (n_lines == 0)
    return RED_B200_OK;
  if (!d_text || !d_starts || !d_lens || !d_out_starts || !d_out_lens || !d_counts ||
      (repl_len && !repl))
    return fail(RED_B200_EAPI, "null buffer");
  DeviceGuard guard(prog->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t *dRepl = nullptr;
  uint64_t *dTotal = nullptr;
  auto done = [&](int rc) {
    scratch_free(dRepl, st);
    scratch_free(dTotal, st);
    return rc;
  };
  cudaError_t e;
  if ((e = scratch_alloc(&dRepl, repl_len + 1, st)) != cudaSuccess ||
      (e = scratch_alloc(&dTotal, 8, st)) != cudaSuccess ||
      (repl_len && (e = cudaMemcpyAsync(dRepl, repl, repl_len, cudaMemcpyHostToDevice, st)) != cudaSuccess))
    return done(cuda_fail(e, "replace staging"));
  int launches = 0;
  const int delim = (out_delim < 0) ? -1 : out_delim;
  // pass 0: substitutions and rewritten length per line
  e = launch_batch_replace(prog->dev, prog->smemTable, false, d_text, d_starts, d_lens, n_lines,
                           style, max_per_line, dRepl, repl_len, delim, d_counts, d_out_lens,
                           nullptr, nullptr, prog->geo, st, &launches);
  if (e != cudaSuccess)
    return done(cuda_fail(e, "replace kernel launch"));
  std::string why;
  int rc = exclusive_offsets(d_out_lens, n_lines, delim >= 0 ? 1u : 0u, d_out_starts, dTotal, st, why);
  if (rc != RED_B200_OK)
    return done(fail(rc, why));
  launches += 3;
  uint64_t total = 0;
  if ((e = cudaMemcpyAsync(&total, dTotal, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
      (e = cudaStreamSynchronize(st)) != cudaSuccess)
    return done(cuda_fail(e, "replace sizes"));
  *out_bytes = total;
  prog->launches += launches;
  if (total > out_capacity)
    return done(fail(RED_B200_ELIMIT, "output buffer too small for the rewritten text"));
  if (total && !d_out_text)
    return done(fail(RED_B200_EAPI, "null output buffer"));
  // pass 1: write
  launches = 0;
  e = launch_batch_replace(prog->dev, prog->smemTable, true, d_text, d_starts, d_lens, n_lines,
                           style, max_per_line, dRepl, repl_len, delim, nullptr, nullptr,
                           d_out_starts, d_out_text, prog->geo, st, &launches);
  if (e != cudaSuccess)
    return done(cuda_fail(e, "replace kernel launch"));
  prog->launches += launches;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess)  // dRepl is pageable-staged: keep it simple
    return done(cuda_fail(e, "replace"));
  return done(RED_B200_OK);
}

namespace {

// host lines -> device (simple staging for the moderate-batch entry points)
struct StagedLines {
  uint8_t *dText = nullptr;
  uint64_t *dStarts = nullptr;
  uint32_t *dLens = nullptr;
  cudaError_t put(const uint8_t *text, uint64_t textLen, const uint64_t *starts,
                  const uint32_t *lens, uint64_t n) {
    cudaError_t e;
    if ((e = cudaMalloc(&dText, textLen + 64)) != cudaSuccess ||
        (e = cudaMalloc(&dStarts, n * 8 + 8)) != cudaSuccess ||
        (e = cudaMalloc(&dLens, n * 4 + 4)) != cudaSuccess ||
        (e = cudaMemcpy(dText, text, textLen, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dStarts, starts, n * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dLens, lens, n * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
      return e;
    return cudaSuccess;
  }
  ~StagedLines() {
    cudaFree(dText);
    cudaFree(dStarts);
    cudaFree(dLens);
  }
};

} // anonymous

int red_b200_replace_batch_host(const red_b200_program *cprog,
                                const uint8_t *text, uint64_t text_len,
                                const uint64_t *starts, const uint32_t *lens,
                                uint64_t n_lines, const uint8_t *repl, uint32_t repl_len,
                                uint64_t max_per_line, int style, int out_delim,
                                uint8_t *out_text, uint64_t out_capacity,
                                uint64_t *out_starts, uint32_t *out_lens,
                                uint32_t *counts, uint64_t *out_bytes) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog || !out_bytes)
    return fail(RED_B200_EAPI, "null argument");
  *out_bytes = 0;
  if (n_lines == 0)
    return RED_B200_OK;
  if (!text || !starts || !lens || !out_starts || !out_lens || !counts)
    return fail(RED_B200_EAPI, "null buffer");
  for (uint64_t i = 0; i < n_lines; ++i)
    if (starts[i] + lens[i] > text_len)
      return fail(RED_B200_EAPI, "line extends past end of text");
  DeviceGuard guard(prog->device);
  StagedLines in;
  uint8_t *dOut = nullptr;
  uint64_t *dOutStarts = nullptr;
  uint32_t *dOutLens = nullptr, *dCounts = nullptr;
  auto done = [&](int rc) {
    cudaFree(dOut); cudaFree(dOutStarts); cudaFree(dOutLens); cudaFree(dCounts);
    return rc;
  };
  cudaError_t e;
  if ((e = in.put(text, text_len, starts, lens, n_lines)) != cudaSuccess ||
      (e = cudaMalloc(&dOut, out_capacity + 1)) != cudaSuccess ||
      (e = cudaMalloc(&dOutStarts, n_lines * 8)) != cudaSuccess ||
      (e = cudaMalloc(&dOutLens, n_lines * 4)) != cudaSuccess ||
      (e = cudaMalloc(&dCounts, n_lines * 4)) != cudaSuccess)
    return done(cuda_fail(e, "replace staging"));
  int rc = red_b200_replace_batch_device(prog, in.dText, in.dStarts, in.dLens, n_lines, repl,
                                         repl_len, max_per_line, style, out_delim, dOut,
                                         out_capacity, dOutStarts, dOutLens, dCounts, out_bytes,
                                         nullptr);
  if ((rc == RED_B200_OK) || (rc == RED_B200_ELIMIT)) {
    // sizes and counts are valid in both cases
    if ((e = cudaMemcpy(out_lens, dOutLens, n_lines * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
        (e = cudaMemcpy(out_starts, dOutStarts, n_lines * 8, cudaMemcpyDeviceToHost)) != cudaSuccess ||
        (e = cudaMemcpy(counts, dCounts, n_lines * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
        ((rc == RED_B200_OK) && *out_bytes &&
         (e = cudaMemcpy(out_text, dOut, *out_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess))
      return done(cuda_fail(e, "replace copy-out"));
  }
  return done(rc);
}

int red_b200_replace(const red_b200_program *prog, const void *text, size_t len,
                     const void *repl, size_t repl_len, uint64_t max, int style,
                     void *out, size_t out_capacity, size_t *out_len, uint64_t *count) {
  if (!out_len || !count)
    return fail(RED_B200_EAPI, "null argument");
  *out_len = 0;
  *count = 0;
  if (len == 0)
    return prog ? RED_B200_OK : fail(RED_B200_EAPI, "null program");
  if (len > 0xffffffffull)
    return fail(RED_B200_ELIMIT, "text too long for replace (4 GB)");
  if (repl_len > 0xffffffffull)
    return fail(RED_B200_ELIMIT, "replacement too long");
  uint64_t start = 0, outStart = 0, bytes = 0;
  uint32_t l = static_cast<uint32_t>(len), outLen = 0, cnt = 0;
  int rc = red_b200_replace_batch_host(prog, static_cast<const uint8_t *>(text), len, &start, &l, 1,
                                       static_cast<const uint8_t *>(repl),
                                       static_cast<uint32_t>(repl_len), max, style, -1,
                                       static_cast<uint8_t *>(out), out_capacity, &outStart,
                                       &outLen, &cnt, &bytes);
  *out_len = bytes;
  *count = cnt;
  return rc;
}

int red_b200_advance_batch_device(const red_b200_program *cprog,
                                  const uint8_t *d_text, const uint64_t *d_starts,
                                  const uint32_t *d_lens, uint64_t n_streams,
                                  uint32_t *d_states, int32_t *d_results,
                                  int32_t *d_hit_results, uint32_t *d_hit_ends,
                                  void *stream) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog)
    return fail(RED_B200_EAPI, "null program");
  if (n_streams == 0)
    return RED_B200_OK;
  if (!d_starts || !d_lens || !d_states || !d_results)
    return fail(RED_B200_EAPI, "null buffer");
  DeviceGuard guard(prog->device);
  int launches = 0;
  cudaError_t err = launch_batch_advance(prog->dev, prog->smemTable, d_text, d_starts, d_lens,
                                         n_streams, d_states, d_results, d_hit_results,
                                         d_hit_ends, prog->geo,
                                         static_cast<cudaStream_t>(stream), &launches);
  if (err != cudaSuccess)
    return cuda_fail(err, "advance kernel launch");
  prog->launches += launches;
  return RED_B200_OK;
}

int red_b200_advance_batch_host(const red_b200_program *cprog,
                                const uint8_t *text, uint64_t text_len,
                                const uint64_t *starts, const uint32_t *lens,
                                uint64_t n_streams, uint32_t *states,
                                int32_t *results, int32_t *hit_results,
                                uint32_t *hit_ends) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog)
    return fail(RED_B200_EAPI, "null program");
  if (n_streams == 0)
    return RED_B200_OK;
  if (!starts || !lens || !states || !results || (text_len && !text))
    return fail(RED_B200_EAPI, "null buffer");
  for (uint64_t i = 0; i < n_streams; ++i)
    if (starts[i] + lens[i] > text_len)
      return fail(RED_B200_EAPI, "segment extends past end of text");
  DeviceGuard guard(prog->device);
  StagedLines in;
  uint32_t *dStates = nullptr, *dHitEnds = nullptr;
  int32_t *dRes = nullptr, *dHit = nullptr;
  auto done = [&](int rc) {
    cudaFree(dStates); cudaFree(dHitEnds); cudaFree(dRes); cudaFree(dHit);
    return rc;
  };
  cudaError_t e;
  static const uint8_t none = 0;
  if ((e = in.put(text ? text : &none, text_len, starts, lens, n_streams)) != cudaSuccess ||
      (e = cudaMalloc(&dStates, n_streams * 4)) != cudaSuccess ||
      (e = cudaMalloc(&dRes, n_streams * 4)) != cudaSuccess ||
      (e = cudaMalloc(&dHit, n_streams * 4)) != cudaSuccess ||
      (e = cudaMalloc(&dHitEnds, n_streams * 4)) != cudaSuccess ||
      (e = cudaMemcpy(dStates, states, n_streams * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
    return done(cuda_fail(e, "advance staging"));
  int rc = red_b200_advance_batch_device(prog, in.dText, in.dStarts, in.dLens, n_streams, dStates,
                                         dRes, dHit, dHitEnds, nullptr);
  if (rc == RED_B200_OK) {
    if ((e = cudaMemcpy(states, dStates, n_streams * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
        (e = cudaMemcpy(results, dRes, n_streams * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
        (hit_results && (e = cudaMemcpy(hit_results, dHit, n_streams * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) ||
        (hit_ends && (e = cudaMemcpy(hit_ends, dHitEnds, n_streams * 4, cudaMemcpyDeviceToHost)) != cudaSuccess))
      rc = cuda_fail(e, "advance copy-out");
  }
  return done(rc);
}

int red_b200_split_lines_device(const uint8_t *d_text, uint64_t text_len,
                                uint8_t delim, uint64_t capacity,
                                uint64_t *d_starts, uint32_t *d_lens,
                                uint64_t *n_lines_out, int device,
                                void *stream) {
  if (!n_lines_out)
    return fail(RED_B200_EAPI, "null argument");
  if (text_len && (!d_text || !d_starts || !d_lens))
    return fail(RED_B200_EAPI, "null buffer");
  if (!usable_device(device))
    return fail(RED_B200_ENODEVICE, "device is not an sm_100 GPU; red_b200 has no CPU fallback");
  DeviceGuard guard(device);
  std::string why;
  int rc = split_lines(d_text, text_len, delim, capacity, d_starts, d_lens,
                       n_lines_out, static_cast<cudaStream_t>(stream), why);
  if (rc != RED_B200_OK)
    return fail(rc, why);
  return RED_B200_OK;
}

int red_b200_compact_matches_device(const int32_t *d_results, uint64_t n_lines,
                                    uint64_t *d_indices, uint64_t *d_count,
                                    int device, void *stream) {
  if (!d_count || (n_lines && (!d_results || !d_indices)))
    return fail(RED_B200_EAPI, "null buffer");
  if (!usable_device(device))
    return fail(RED_B200_ENODEVICE, "device is not an sm_100 GPU; red_b200 has no CPU fallback");
  DeviceGuard guard(device);
  std::string why;
  int rc = compact_matches(d_results, n_lines, d_indices, d_count,
                           static_cast<cudaStream_t>(stream), why);
  if (rc != RED_B200_OK)
    return fail(rc, why);
  return RED_B200_OK;
}

int red_b200_match_stream_device(const red_b200_program *cprog,
                                 const uint8_t *d_text, uint64_t len,
                                 int verb, int style, red_b200_outcome *out,
                                 void *stream) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  if (!prog || !out)
    return fail(RED_B200_EAPI, "null argument");
  if (!valid_style(style))
    return fail(RED_B200_EEXEC, "unsupported style");
  if ((verb != RED_B200_VERB_CHECK) && (verb != RED_B200_VERB_MATCH))
    return fail(RED_B200_EEXEC, "stream path supports check and match");
  if (len && !d_text)
    return fail(RED_B200_EAPI, "null text");
  DeviceGuard guard(prog->device);
  std::string why;
  int launches = 0;
  /* The main walk runs on the hot image (stream_hot_kernel).  A program that
     has seen no batch yet has no profile: take one from this text - pieces of
     256 bytes spread over it, each walked from the initial state (what a DFA
     that forgets enters behind an arbitrary cut is what it enters anywhere). */
  HotDev hot;
  bool haveHot = false;
  if (prog->hotEnabled) {
    if (prog->hotProfileWanted && !prog->hotProfiled.load() && (len >= (1u << 20))) {
      const uint64_t pieces = 32768, pieceLen = 256;
      const uint64_t step = (len - pieceLen) / pieces;
      std::vector<uint64_t> hs(pieces);
      std::vector<uint32_t> hl(pieces, static_cast<uint32_t>(pieceLen));
      for (uint64_t i = 0; i < pieces; ++i)
        hs[i] = i * step;
      uint64_t *dS = nullptr;
      uint32_t *dL = nullptr;
      cudaStream_t st = static_cast<cudaStream_t>(stream);
      if ((cudaMalloc(&dS, pieces * 8) == cudaSuccess) && (cudaMalloc(&dL, pieces * 4) == cudaSuccess) &&
          (cudaMemcpyAsync(dS, hs.data(), pieces * 8, cudaMemcpyHostToDevice, st) == cudaSuccess) &&
          (cudaMemcpyAsync(dL, hl.data(), pieces * 4, cudaMemcpyHostToDevice, st) == cudaSuccess) &&
          (cudaStreamSynchronize(st) == cudaSuccess)) {
        BatchArgs a{};
        a.text = d_text;
        a.starts = dS;
        a.lens = dL;
        a.nLines = pieces;
        profile_hot(prog, a, st, false);
      }
      else
        cudaGetLastError();
      cudaFree(dS);
      cudaFree(dL);
    }
    std::lock_guard<std::mutex> lock(prog->hotMu);
    hot = prog->hotStream;
    haveHot = true;
  }
  int rc = match_stream(prog->dev, prog->smemTable, prog->img, haveHot ? &hot : nullptr, prog->geo,
                        d_text, len, style, out,
                        static_cast<cudaStream_t>(stream), &launches, why);
  prog->launches += launches;
  if (rc != RED_B200_OK)
    return fail(rc, why);
  if (verb == RED_B200_VERB_CHECK) {
    out->start = 0;
    out->end = 0;
  }
  return RED_B200_OK;
}

int red_b200_compile(int n, const char *const *patterns,
                     const uint64_t *pattern_lens, const int32_t *results,
                     const uint32_t *flags, const int32_t *langs, int format,
                     uint64_t max_states, void **out, uint64_t *out_len) {
  if (!out || !out_len || (n < 0) ||
      (n && (!patterns || !pattern_lens || !results || !flags || !langs)))
    return fail(RED_B200_EAPI, "null argument");
  std::string buf, why;
  int rc = compile_patterns(n, patterns, pattern_lens, results, flags, langs,
                            format, max_states, buf, why);
  if (rc != RED_B200_OK)
    return fail(rc, why);
  void *mem = malloc(buf.size() ? buf.size() : 1);
  if (!mem)
    return fail(RED_B200_ELIMIT, "out of memory");
  memcpy(mem, buf.data(), buf.size());
  *out = mem;
  *out_len = buf.size();
  return RED_B200_OK;
}

void red_b200_free(void *p) { free(p); }

} // extern "C"
