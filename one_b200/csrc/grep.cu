/* grep.cu - the flume/loggle workload end to end on the GPU:

       log blob  ->  line index  ->  match every line  ->  matching lines

   red_b200_grep_host() takes a host buffer of delimiter-separated lines and
   returns, in text order, one record per line whose result is > 0.  Only the
   text crosses PCIe on the way in and only the matching lines' records on the
   way out; line index, matching (red_b200's per-line kernel, verb match) and
   compaction stay in HBM.  Per line the outcome is exactly what
   match(exec, line, style) of the reference gives (include/Matcher.h:79).

   The blob is cut into pieces that end on a delimiter; three slots keep the
   copy-in of one piece, the kernels of another and the copy-out of a third in
   flight. */

#include <algorithm>
#include <functional>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "program_handle.h"
#include "stream.h"

using namespace red_b200;

namespace {

thread_local std::string g_why;

constexpr int kSlots = 3;

__global__ void gather_matches_kernel(const uint64_t *indices, const uint64_t *count,
                                      uint64_t capacity, const uint64_t *starts,
                                      const uint32_t *lens, const int32_t *results,
                                      const uint32_t *mstarts, const uint32_t *mends,
                                      uint64_t textBase, red_b200_line_match *out) {
  const uint64_t n = min(*count, capacity);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       i < n; i += stride) {
    const uint64_t line = indices[i];
    red_b200_line_match m;
    m.line_start = textBase + starts[line];
    m.line_len = lens[line];
    m.result = results[line];
    m.match_start = mstarts[line];
    m.match_end = mends[line];
    out[i] = m;
  }
}

} // anonymous

struct GrepPipe {
  size_t textCap = 0, lineCap = 0;
  cudaStream_t stream[kSlots]{};
  cudaEvent_t done[kSlots]{};
  uint8_t *dText[kSlots]{};
  uint64_t *dStarts[kSlots]{};
  uint32_t *dLens[kSlots]{};
  int32_t *dRes[kSlots]{};
  uint32_t *dMs[kSlots]{};
  uint32_t *dMe[kSlots]{};
  uint64_t *dIdx[kSlots]{};
  uint64_t *dCount[kSlots]{};
  LineCount *dInfo[kSlots]{};
  red_b200_line_match *dOut[kSlots]{};
  SplitScratch scratch[kSlots];
  // pinned: per slot {matches, LineCount}
  uint64_t *hCount[kSlots]{};
  LineCount *hInfo[kSlots]{};

  void release() {
    for (int s = 0; s < kSlots; ++s) {
      if (stream[s]) cudaStreamDestroy(stream[s]);
      if (done[s]) cudaEventDestroy(done[s]);
      cudaFree(dText[s]); cudaFree(dStarts[s]); cudaFree(dLens[s]); cudaFree(dRes[s]);
      cudaFree(dMs[s]); cudaFree(dMe[s]); cudaFree(dIdx[s]); cudaFree(dCount[s]);
      cudaFree(dInfo[s]); cudaFree(dOut[s]);
      scratch[s].release();
      if (hCount[s]) cudaFreeHost(hCount[s]);
      if (hInfo[s]) cudaFreeHost(hInfo[s]);
    }
  }
};

namespace red_b200 {
void release_grep_pipe(red_b200_program *prog) {
  if (prog->grep) {
    prog->grep->release();
    delete prog->grep;
    prog->grep = nullptr;
  }
}
}

namespace {

int cuda_fail(cudaError_t e, const char *what) {
  g_why = std::string(what) + ": " + cudaGetErrorString(e);
  return RED_B200_ECUDA;
}

#define CUDA_TRY(expr)                    \
  do {                                    \
    cudaError_t e_ = (expr);              \
    if (e_ != cudaSuccess)                \
      return cuda_fail(e_, #expr);        \
  } while (0)

int ensure_grep(red_b200_program *prog, size_t textCap, size_t lineCap) {
  GrepPipe *gp = prog->grep;
  if (gp && (gp->textCap >= textCap) && (gp->lineCap >= lineCap))
    return RED_B200_OK;
  release_grep_pipe(prog);
  gp = new GrepPipe;
  gp->textCap = textCap;
  gp->lineCap = lineCap;
  prog->grep = gp;
  for (int s = 0; s < kSlots; ++s) {
    CUDA_TRY(cudaStreamCreateWithFlags(&gp->stream[s], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&gp->done[s], cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&gp->dText[s], textCap + 64));
    CUDA_TRY(cudaMalloc(&gp->dStarts[s], (lineCap + 1) * 8));
    CUDA_TRY(cudaMalloc(&gp->dLens[s], (lineCap + 1) * 4));
    CUDA_TRY(cudaMalloc(&gp->dRes[s], lineCap * 4));
    CUDA_TRY(cudaMalloc(&gp->dMs[s], lineCap * 4));
    CUDA_TRY(cudaMalloc(&gp->dMe[s], lineCap * 4));
    CUDA_TRY(cudaMalloc(&gp->dIdx[s], lineCap * 8));
    CUDA_TRY(cudaMalloc(&gp->dCount[s], 8));
    CUDA_TRY(cudaMalloc(&gp->dInfo[s], sizeof(LineCount)));
    CUDA_TRY(cudaMalloc(&gp->dOut[s], lineCap * sizeof(red_b200_line_match)));
    CUDA_TRY(gp->scratch[s].alloc(textCap + 64));
    CUDA_TRY(cudaMallocHost(&gp->hCount[s], 8));
    CUDA_TRY(cudaMallocHost(&gp->hInfo[s], sizeof(LineCount)));
  }
  return RED_B200_OK;
}

} // anonymous

extern "C" int red_b200_grep_host(const red_b200_program *cprog, const uint8_t *text,
                                  uint64_t text_len, uint8_t delim, int style,
                                  red_b200_line_match *matches, uint64_t capacity,
                                  uint64_t *n_matches, uint64_t *n_lines) {
  red_b200_program *prog = const_cast<red_b200_program *>(cprog);
  auto fail = [&](int code, const std::string &msg) {
    set_last_error(msg);
    return code;
  };
  if (!prog || !n_matches || !n_lines)
    return fail(RED_B200_EAPI, "null argument");
  *n_matches = 0;
  *n_lines = 0;
  if ((style < 1) || (style > 5))
    return fail(RED_B200_EEXEC, "unsupported style");
  if (text_len == 0)
    return RED_B200_OK;
  if (!text || (capacity && !matches))
    return fail(RED_B200_EAPI, "null buffer");

  std::lock_guard<std::mutex> lock(prog->mu);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(prog->device);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};

  size_t piece = 32u << 20;
  if (const char *env = getenv("RED_B200_HOST_CHUNK_BYTES"))
    piece = std::max<size_t>(1 << 12, strtoull(env, nullptr, 0));
  // The index of a piece holds one line per 8 bytes of text.  A piece that has
  // more (word lists, runs of blank lines) is taken again in smaller pieces,
  // sized by the line count the split reported - down to pieces of lineCap
  // bytes, which cannot overflow whatever they hold.
  const size_t lineCap = piece / 8 + 16;
  int rc = ensure_grep(prog, piece, lineCap);
  if (rc != RED_B200_OK)
    return fail(rc, g_why);
  GrepPipe *gp = prog->grep;

  struct InFlight { bool busy = false; uint64_t base = 0, bytes = 0; } slot[kSlots];
  uint64_t written = 0, lines = 0;
  bool truncated = false;
  std::string why;

  // text [pos, end) through slot s: copy in, index, match, compact, gather
  auto submit = [&](int s, uint64_t pos, uint64_t end) -> int {
    cudaStream_t st = gp->stream[s];
    const uint64_t n = end - pos;
    CUDA_TRY(cudaMemcpyAsync(gp->dText[s], text + pos, n, cudaMemcpyHostToDevice, st));
    int r = split_lines_async(gp->dText[s], n, delim, gp->lineCap, gp->dStarts[s], gp->dLens[s],
                              gp->dInfo[s], gp->scratch[s], st, why);
    if (r != RED_B200_OK) {
      g_why = why;
      return r;
    }
    BatchArgs args;
    args.text = gp->dText[s];
    args.starts = gp->dStarts[s];
    args.lens = gp->dLens[s];
    args.nLines = gp->lineCap;
    args.nLinesDev = &gp->dInfo[s]->nLines;
    args.results = gp->dRes[s];
    args.matchStarts = gp->dMs[s];
    args.matchEnds = gp->dMe[s];
    r = launch_batch_for(prog, args, RED_B200_VERB_MATCH, style, st, why);
    if (r != RED_B200_OK) {
      g_why = why;
      return r;
    }
    r = compact_matches_n(gp->dRes[s], gp->lineCap, &gp->dInfo[s]->nLines, gp->dIdx[s],
                          gp->dCount[s], gp->scratch[s], st, why);
    if (r != RED_B200_OK) {
      g_why = why;
      return r;
    }
    gather_matches_kernel<<<148 * 2, 256, 0, st>>>(gp->dIdx[s], gp->dCount[s], gp->lineCap,
                                                   gp->dStarts[s], gp->dLens[s], gp->dRes[s],
                                                   gp->dMs[s], gp->dMe[s], pos, gp->dOut[s]);
    prog->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(gp->hCount[s], gp->dCount[s], 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(gp->hInfo[s], gp->dInfo[s], sizeof(LineCount), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(gp->done[s], st));
    slot[s].busy = true;
    slot[s].base = pos;
    slot[s].bytes = n;
    return RED_B200_OK;
  };

  // copy out what a finished slot produced; a piece whose index overflowed is
  // taken again, in smaller pieces, before anything behind it is looked at
  std::function<int(int)> drain = [&](int s) -> int {
    if (!slot[s].busy)
      return RED_B200_OK;
    CUDA_TRY(cudaEventSynchronize(gp->done[s]));
    slot[s].busy = false;
    const LineCount info = *gp->hInfo[s];
    if (info.overflow) {
      const uint64_t base = slot[s].base, end = base + slot[s].bytes;
      // half of what the average density would allow, never more than half the
      // piece that failed; lineCap bytes always fit
      uint64_t sub = slot[s].bytes / std::max<uint64_t>(2, 2 * ((info.wanted + gp->lineCap - 1) / gp->lineCap));
      sub = std::max<uint64_t>(sub, std::min<uint64_t>(gp->lineCap, slot[s].bytes / 2));
      sub = std::max<uint64_t>(sub, 16);
      for (uint64_t p = base; p < end;) {
        uint64_t q = std::min(end, p + sub);
        if (q < end) {
          // cut just behind a delimiter; a line longer than `sub` is kept whole
          const void *d = memrchr(text + p, delim, q - p);
          if (!d)
            d = memchr(text + q, delim, end - q);
          q = d ? static_cast<uint64_t>(static_cast<const uint8_t *>(d) - text) + 1 : end;
        }
        int r = submit(s, p, q);
        if (r != RED_B200_OK)
          return r;
        if ((r = drain(s)) != RED_B200_OK)
          return r;
        p = q;
      }
      return RED_B200_OK;
    }
    lines += info.nLines;
    uint64_t cnt = *gp->hCount[s];
    uint64_t room = capacity - std::min(capacity, written);
    if (cnt > room) {
      truncated = true;
      cnt = room;
    }
    if (cnt) {
      CUDA_TRY(cudaMemcpyAsync(matches + written, gp->dOut[s], cnt * sizeof(red_b200_line_match),
                               cudaMemcpyDeviceToHost, gp->stream[s]));
      CUDA_TRY(cudaStreamSynchronize(gp->stream[s]));
    }
    written += cnt;
    return RED_B200_OK;
  };

  uint64_t pos = 0;
  size_t ci = 0;
  for (; pos < text_len; ++ci) {
    // a piece ends just after the last delimiter inside [pos, pos + piece), or
    // at the end of the text; a single line longer than a piece is refused
    uint64_t end = std::min<uint64_t>(text_len, pos + piece);
    if (end < text_len) {
      const void *p = memrchr(text + pos, delim, end - pos);
      if (!p)
        return fail(RED_B200_ELIMIT, "a line is longer than the staging piece (RED_B200_HOST_CHUNK_BYTES)");
      end = static_cast<const uint8_t *>(p) - text + 1;
    }
    const int s = static_cast<int>(ci % kSlots);
    if ((rc = drain(s)) != RED_B200_OK)
      return fail(rc, g_why);
    if ((rc = submit(s, pos, end)) != RED_B200_OK)
      return fail(rc, g_why);
    pos = end;
    // keep text order: the oldest slot is always drained first (round robin)
  }
  // drain what is still in flight, in text order (slots were filled round robin)
  for (size_t k = (ci > static_cast<size_t>(kSlots)) ? ci - kSlots : 0; k < ci; ++k)
    if ((rc = drain(static_cast<int>(k % kSlots))) != RED_B200_OK)
      return fail(rc, g_why);
  *n_matches = written;
  *n_lines = lines;
  if (truncated)
    return fail(RED_B200_ELIMIT, "more matching lines than the output holds");
  return RED_B200_OK;
}
