/* red_b200.h - C ABI of the B200-native RED matcher (libred_b200.so).

   This is the drop-in boundary for RED's data-parallel hot path: executing a
   compiled, minimized, serialized DFA (the reference's `Executable`) over
   batches of byte strings and over long byte streams, on sm_100a.

   Everything here is plain C: pointers, sizes, ints.  No C++/torch types.
   "device pointer" means a pointer into the HBM of the program's GPU (e.g. a
   cudaMalloc'd buffer or a torch tensor's data_ptr()); "stream" is a
   cudaStream_t passed as void* (NULL = the default stream).

   Reference interfaces each entry point stands in for
   (paths relative to /root/reference/quol/red):

     red_b200_program_upload      Executable(const CopyTag&, string_view)
                                    include/Executable.h:47, validate() at
                                    lib/Executable.cpp:159, checkHeader() at
                                    lib/Serializer.cpp:270
     red_b200_match               match(exec, ptr, len, style) and
                                  check/scan/search siblings
                                    include/Matcher.h:74-94
     red_b200_match_batch_*       the per-line loop over the same calls that a
                                  RED user writes (tools/thr_red.cpp,
                                  tools/big_red.cpp); no batch call exists in
                                  the reference
     red_b200_match_stream_*      match(exec, ptr, len, style) on one very long
                                  text  include/Matcher.h:79
     red_b200_compile             Parser::addAs + compileToSerialized
                                    include/Parser.h:98, include/Compile.h:30

   Style and verb numbers are the reference's (include/Matcher.h:62-69).
   All functions return 0 on success or a negative RED_B200_E* code; the text
   of the last error on the calling thread is red_b200_last_error().
   There is no CPU fallback: without a usable sm_100 device every compute entry
   point fails with RED_B200_ENODEVICE.
 */

#ifndef RED_B200_H
#define RED_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RED_B200_ABI_VERSION 1

/* match styles: Matcher.h:62-69 */
enum {
  RED_B200_STY_INSTANT = 1,
  RED_B200_STY_FIRST   = 2,
  RED_B200_STY_TANGENT = 3,
  RED_B200_STY_LAST    = 4,
  RED_B200_STY_FULL    = 5
};

/* verbs: Matcher.h:9-13.  check = match without positions;
   scan = search without positions. */
enum {
  RED_B200_VERB_CHECK  = 0,
  RED_B200_VERB_MATCH  = 1,
  RED_B200_VERB_SCAN   = 2,
  RED_B200_VERB_SEARCH = 3
};

/* error codes; the class mirrors the reference's exceptions (Except.h) */
enum {
  RED_B200_OK         =  0,
  RED_B200_EAPI       = -1,  /* RedExceptApi: bad argument / bad program   */
  RED_B200_EEXEC      = -2,  /* RedExceptExec: unsupported style or verb   */
  RED_B200_ELIMIT     = -3,  /* RedExceptLimit: too big                    */
  RED_B200_EPARSE     = -4,  /* RedExceptParse: malformed regex            */
  RED_B200_ECUDA      = -5,  /* a CUDA call failed                         */
  RED_B200_ENODEVICE  = -6,  /* no sm_100 GPU visible: no CPU fallback     */
  RED_B200_EINTERNAL  = -7
};

/* Outcome: include/Outcome.h:33-37 */
typedef struct {
  int32_t  result;
  uint64_t start;
  uint64_t end;
} red_b200_outcome;

typedef struct red_b200_program red_b200_program;

/* where the kernels keep the transition table */
enum {
  RED_B200_TABLE_SMEM_U16   = 1, /* shared memory, 2-byte entries            */
  RED_B200_TABLE_GLOBAL_U32 = 2  /* L2-resident global table, 4-byte entries */
};

typedef struct {
  uint32_t abi_version;
  uint32_t format;        /* serialized value width: 1, 2 or 4              */
  uint32_t state_count;
  uint32_t class_count;   /* byte equivalence classes = maxChar + 1         */
  uint32_t leader_len;
  uint32_t table_mode;    /* RED_B200_TABLE_*                               */
  uint32_t row_stride;    /* entries per row in the device table            */
  uint32_t accepting_states;
  uint32_t pure_dead_states;
  uint32_t max_result;
  uint64_t table_bytes;   /* device transition table size                   */
  int32_t  device;
  int32_t  sm_count;
} red_b200_program_info;

const char *red_b200_last_error(void);
int red_b200_abi_version(void);

/* number of usable (compute capability 10.x) devices; 0 if none */
int red_b200_device_count(void);

/* Validate a serialized DFA exactly as the reference does (magic, version,
   FNV-1a checksum, format) plus bounds checks the GPU needs, re-lay it out for
   the kernels and upload it to `device`.  The serialized bytes are not kept. */
int red_b200_program_upload(const void *serialized, size_t len, int device,
                            red_b200_program **out);
void red_b200_program_free(red_b200_program *prog);
int red_b200_program_get_info(const red_b200_program *prog,
                              red_b200_program_info *info);

/* Kernels launched through this program so far, and how many of them were
   launches of the staged batch kernel (diagnostics; bench.py's gpu_launches). */
long red_b200_program_launch_count(const red_b200_program *prog);
long red_b200_program_hot_launch_count(const red_b200_program *prog);
/* With RED_B200_HOT_DEBUG=1 in the environment at upload: cycle counters of
   the staged batch kernel since the last call (tools/hot_sweep.py prints
   them); all zero otherwise.  Synchronises the device. */
int red_b200_program_debug_counters(red_b200_program *prog, uint64_t out[16]);

/* Host-only validation; returns NULL when good, else the reference's wording
   of the complaint (checkHeader, lib/Serializer.cpp:270). */
const char *red_b200_check_header(const void *serialized, size_t len);

/* ---- batch of byte strings (lines) ------------------------------------ */

/* Line i is text[starts[i] .. starts[i]+lens[i]).  Per line the result id,
   and (verbs match/search) where the match started/ended, exactly as
   Outcome would hold them; results of 0 come with start = end = 0.
   match_starts / match_ends may be NULL when not wanted.
   All pointers are device pointers; the call is asynchronous on `stream`.
   The kernels read text in aligned 16-byte blocks: the text allocation must be
   readable from the 16-byte boundary at or below the first line's start to the
   16-byte boundary at or above the last line's end (any cudaMalloc'd or torch
   buffer is, CUDA allocations being 256-byte granular). */
int red_b200_match_batch_device(const red_b200_program *prog,
                                const uint8_t *d_text,
                                const uint64_t *d_starts,
                                const uint32_t *d_lens,
                                uint64_t n_lines, int verb, int style,
                                int32_t *d_results,
                                uint32_t *d_match_starts,
                                uint32_t *d_match_ends,
                                void *stream);

/* Same with host buffers: text and index go to the GPU, results come back, in
   chunks through three slots so that the copies of one chunk overlap the
   kernels of another.  That overlap needs PINNED host buffers (cudaHostAlloc /
   cudaHostRegister, torch pin_memory): from pageable memory the CUDA runtime
   stages every copy itself and a device-to-host copy returns only when it is
   done, so the chunks then run one after the other - correct, but at about the
   speed of the slowest copy.  Synchronous; on an error nothing is left in
   flight. */
int red_b200_match_batch_host(const red_b200_program *prog,
                              const uint8_t *text, uint64_t text_len,
                              const uint64_t *starts, const uint32_t *lens,
                              uint64_t n_lines, int verb, int style,
                              int32_t *results, uint32_t *match_starts,
                              uint32_t *match_ends);

/* One text, host memory: the reference's match()/check()/scan()/search().
   Long texts take the chunked stream path. */
int red_b200_match(const red_b200_program *prog, const void *text, size_t len,
                   int verb, int style, red_b200_outcome *out);

/* ---- per line, all matches: matchAll() and Red::collect() ----------------- */

typedef struct {
  int32_t  result;
  uint32_t start;   /* relative to the line, as Outcome's are to the text */
  uint32_t end;
} red_b200_span;

enum { RED_B200_MULTI_MATCH_ALL = 0,  /* matchAll(), Matcher.h:127             */
       RED_B200_MULTI_COLLECT   = 1   /* Red::collect(), Red.h:109             */ };

/* For every line: counts[i] outcomes were found; the first
   min(counts[i], max_per_line) are in spans[i * max_per_line ...].
   Device pointers; asynchronous on `stream`. */
int red_b200_multi_batch_device(const red_b200_program *prog,
                                const uint8_t *d_text, const uint64_t *d_starts,
                                const uint32_t *d_lens, uint64_t n_lines,
                                int which, uint32_t max_per_line,
                                uint32_t *d_counts, red_b200_span *d_spans,
                                void *stream);

/* Same with host buffers.  Synchronous. */
int red_b200_multi_batch_host(const red_b200_program *prog,
                              const uint8_t *text, uint64_t text_len,
                              const uint64_t *starts, const uint32_t *lens,
                              uint64_t n_lines, int which, uint32_t max_per_line,
                              uint32_t *counts, red_b200_span *spans);

/* ---- replace() ------------------------------------------------------------ */

/* Rewrite every line: each non-overlapping match, found by trying every
   position of the line in order, is substituted by `repl`; at most
   max_per_line substitutions per line (reference: replace(),
   include/Matcher.h:100-125; replaceCore, :643-707; `max` as there).
   Output: the rewritten lines packed into d_out_text, each followed by the
   byte out_delim unless out_delim < 0; d_out_starts/d_out_lens index them
   (lengths exclude the delimiter); d_counts[i] = substitutions made in line i.
   `repl` is HOST memory.  *out_bytes (host) gets the size of the packed
   output; if that exceeds out_capacity nothing is written to d_out_text and
   the call returns RED_B200_ELIMIT - call again with a buffer of *out_bytes.
   Device pointers otherwise; synchronous on `stream`. */
int red_b200_replace_batch_device(const red_b200_program *prog,
                                  const uint8_t *d_text, const uint64_t *d_starts,
                                  const uint32_t *d_lens, uint64_t n_lines,
                                  const uint8_t *repl, uint32_t repl_len,
                                  uint64_t max_per_line, int style, int out_delim,
                                  uint8_t *d_out_text, uint64_t out_capacity,
                                  uint64_t *d_out_starts, uint32_t *d_out_lens,
                                  uint32_t *d_counts, uint64_t *out_bytes,
                                  void *stream);

/* Same with host buffers (moderate batches).  Synchronous. */
int red_b200_replace_batch_host(const red_b200_program *prog,
                                const uint8_t *text, uint64_t text_len,
                                const uint64_t *starts, const uint32_t *lens,
                                uint64_t n_lines, const uint8_t *repl, uint32_t repl_len,
                                uint64_t max_per_line, int style, int out_delim,
                                uint8_t *out_text, uint64_t out_capacity,
                                uint64_t *out_starts, uint32_t *out_lens,
                                uint32_t *counts, uint64_t *out_bytes);

/* One text (the reference's replace(exec, ptr, len, repl, out, max, style)):
   writes at most out_capacity bytes, *out_len is the rewritten length
   (RED_B200_ELIMIT if it did not fit), *count the substitutions. */
int red_b200_replace(const red_b200_program *prog, const void *text, size_t len,
                     const void *repl, size_t repl_len, uint64_t max, int style,
                     void *out, size_t out_capacity, size_t *out_len, uint64_t *count);

/* ---- StatefulMatcher -------------------------------------------------------- */

/* The reference's StatefulMatcher (include/Matcher.h:770-794) keeps one DFA
   state and is fed byte by byte.  Here the caller keeps the state - an opaque
   u32 per stream, valid only with the program that produced it - and feeds a
   SEGMENT per stream per call; RED_B200_STATE_FRESH (like any value that is
   not a state of this program) starts a new matcher.
   d_results[i] is result() after the segment (for an empty segment: of the
   state as it is).  Optional: d_hit_results[i] / d_hit_ends[i], the result at
   and the offset one past the last accepting position inside the segment
   (0 / 0 if none) - what a caller polling advance() per byte would have seen
   last.  Device pointers; asynchronous on `stream`. */
#define RED_B200_STATE_FRESH 0xffffffffu
int red_b200_advance_batch_device(const red_b200_program *prog,
                                  const uint8_t *d_text, const uint64_t *d_starts,
                                  const uint32_t *d_lens, uint64_t n_streams,
                                  uint32_t *d_states, int32_t *d_results,
                                  int32_t *d_hit_results, uint32_t *d_hit_ends,
                                  void *stream);

/* Same with host buffers.  Synchronous. */
int red_b200_advance_batch_host(const red_b200_program *prog,
                                const uint8_t *text, uint64_t text_len,
                                const uint64_t *starts, const uint32_t *lens,
                                uint64_t n_streams, uint32_t *states,
                                int32_t *results, int32_t *hit_results,
                                uint32_t *hit_ends);

/* ---- line index and match compaction ----------------------------------- */

/* Split d_text[0..text_len) at `delim` bytes: writes the start offset and the
   length (delimiter excluded) of every line; a trailing piece without a
   delimiter counts as a line.  *n_lines_out gets the count (host memory).
   Capacity is in lines; RED_B200_ELIMIT if it would overflow.  Synchronous on
   `stream` (it has to return the count). */
int red_b200_split_lines_device(const uint8_t *d_text, uint64_t text_len,
                                uint8_t delim, uint64_t capacity,
                                uint64_t *d_starts, uint32_t *d_lens,
                                uint64_t *n_lines_out, int device,
                                void *stream);

/* Warp-ballot compaction of the matching lines: writes the indices of lines
   with result > 0 in increasing order to d_indices (capacity n_lines) and the
   count to *d_count (device, uint64).  Asynchronous on `stream`. */
int red_b200_compact_matches_device(const int32_t *d_results, uint64_t n_lines,
                                    uint64_t *d_indices, uint64_t *d_count,
                                    int device, void *stream);

/* ---- a log blob, end to end (the flume / loggle workload) ---------------- */

typedef struct {
  uint64_t line_start;   /* byte offset of the line in the text              */
  uint32_t line_len;     /* without the delimiter                            */
  int32_t  result;       /* > 0                                              */
  uint32_t match_start;  /* Outcome.start_ / end_, relative to the line      */
  uint32_t match_end;
} red_b200_line_match;

/* text is a sequence of lines separated by `delim` (a last line without one
   counts).  Every line is matched as match(exec, line, style) would
   (include/Matcher.h:79); the lines with a result > 0 are returned in text
   order.  Line index, matching and compaction run on the GPU: only the text
   goes in and only these records come out.  *n_lines gets the number of lines
   seen.  RED_B200_ELIMIT if there are more matching lines than `capacity`
   (the first `capacity` are stored) or if a line is longer than the staging
   piece (32 MiB).  Any line density is fine: a piece with more lines than its
   index holds (one per 8 bytes) is taken again in smaller pieces.
   Synchronous. */
int red_b200_grep_host(const red_b200_program *prog, const uint8_t *text,
                       uint64_t text_len, uint8_t delim, int style,
                       red_b200_line_match *matches, uint64_t capacity,
                       uint64_t *n_matches, uint64_t *n_lines);

/* ---- one long byte stream ---------------------------------------------- */

/* match()/check() of one long device-resident text (reference:
   match(exec, ptr, len, style), include/Matcher.h:79, over a single text).
   The text is taken in windows of growing size and cut into chunks, one lane
   each; a chunk's start state is speculated from a lookback window, every
   chunk is walked in parallel, chunk i is confirmed against the end state of
   chunk i-1 and the unconfirmed ones are walked again from the right state
   until nothing changes: exact for every DFA, one or two rounds for DFAs that
   forget their past.  A walk that is decided (first match for Instant/First/
   Tangent, a state that can never be left) ends with its window; the start
   position is searched for backwards only when an outcome needs it.
   d_text must be 16-byte aligned.  Verbs check and match.  Synchronous on
   `stream`. */
int red_b200_match_stream_device(const red_b200_program *prog,
                                 const uint8_t *d_text, uint64_t len,
                                 int verb, int style, red_b200_outcome *out,
                                 void *stream);

/* ---- host side: regex -> serialized DFA (RED's compile pipeline) -------- */

/* languages: Parser.h:43-49; flags: Consts.h:14-18; formats: Serializer.h:37 */
enum { RED_B200_LANG_RAW = 1, RED_B200_LANG_AUTO = 2, RED_B200_LANG_GLOB = 3,
       RED_B200_LANG_EXACT = 4 };
enum { RED_B200_F_IGNORE_CASE = 1, RED_B200_F_LOOSE_START = 2,
       RED_B200_F_LOOSE_END = 4 };
enum { RED_B200_FMT_1 = 1, RED_B200_FMT_2 = 2, RED_B200_FMT_4 = 4,
       RED_B200_FMT_AUTO = 255 };

/* Parse n patterns, build the NFA, determinize, minimize and serialize into
   the reference's REDA format (loadable by the reference's Executable as
   well).  *out is malloc'd; release with red_b200_free().  max_states = 0
   means no budget (Budget.h). */
int red_b200_compile(int n, const char *const *patterns,
                     const uint64_t *pattern_lens, const int32_t *results,
                     const uint32_t *flags, const int32_t *langs, int format,
                     uint64_t max_states, void **out, uint64_t *out_len);
void red_b200_free(void *p);

#ifdef __cplusplus
}
#endif

#endif /* RED_B200_H */
