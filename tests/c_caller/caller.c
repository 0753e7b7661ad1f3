/* caller.c - a plain C user of libred_b200.so (include/red_b200.h): no Python,
   no C++.  Built and run by tests/test_c_caller.py on the GPU box.

     caller <program.reda> <lines.txt> <expected.txt>

   lines.txt: newline-separated lines.  expected.txt (written by the test from
   the oracle's answers): for every line "result start end" of match/Last, then
   for the first 16 lines "result start end" of match/Instant on the line
   alone, then the length of line 0 rewritten by replace(.., "<X>", Last) and
   its substitution count.  Exit code 0 = everything agreed. */

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "red_b200.h"

static unsigned char *slurp(const char *path, size_t *len) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    perror(path);
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  unsigned char *p = malloc((size_t)n + 64);
  if (fread(p, 1, (size_t)n, f) != (size_t)n) {
    perror("fread");
    exit(2);
  }
  fclose(f);
  *len = (size_t)n;
  return p;
}

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != RED_B200_OK) {                                              \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, red_b200_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  if (argc != 4) {
    fprintf(stderr, "usage: caller program lines expected\n");
    return 2;
  }
  size_t progLen, textLen;
  unsigned char *prog = slurp(argv[1], &progLen);
  unsigned char *text = slurp(argv[2], &textLen);
  FILE *exp = fopen(argv[3], "r");
  if (!exp) {
    perror(argv[3]);
    return 2;
  }
  if (red_b200_abi_version() != RED_B200_ABI_VERSION || red_b200_device_count() < 1) {
    fprintf(stderr, "no usable device / ABI mismatch\n");
    return 1;
  }
  if (red_b200_check_header(prog, progLen) != NULL) {
    fprintf(stderr, "header: %s\n", red_b200_check_header(prog, progLen));
    return 1;
  }

  /* line index */
  uint64_t n = 0;
  for (size_t i = 0; i < textLen; ++i)
    n += text[i] == '\n';
  uint64_t *starts = malloc(n * sizeof *starts);
  uint32_t *lens = malloc(n * sizeof *lens);
  int32_t *res = malloc(n * sizeof *res);
  uint32_t *ms = malloc(n * sizeof *ms), *me = malloc(n * sizeof *me);
  uint64_t k = 0, at = 0;
  for (size_t i = 0; i < textLen; ++i)
    if (text[i] == '\n') {
      starts[k] = at;
      lens[k] = (uint32_t)(i - at);
      ++k;
      at = i + 1;
    }

  red_b200_program *p = NULL;
  CHECK(red_b200_program_upload(prog, progLen, 0, &p));
  red_b200_program_info info;
  CHECK(red_b200_program_get_info(p, &info));

  int bad = 0;
  /* the batch call */
  CHECK(red_b200_match_batch_host(p, text, textLen, starts, lens, n, RED_B200_VERB_MATCH, RED_B200_STY_LAST,
                                  res, ms, me));
  for (uint64_t i = 0; i < n; ++i) {
    long r, s, e;
    if (fscanf(exp, "%ld %ld %ld", &r, &s, &e) != 3)
      return 2;
    if (r != res[i] || s != (long)ms[i] || e != (long)me[i]) {
      if (bad++ < 5)
        fprintf(stderr, "line %llu: got %d %u %u, want %ld %ld %ld\n", (unsigned long long)i, res[i], ms[i], me[i],
                r, s, e);
    }
  }
  /* one text at a time: the reference's match(exec, ptr, len, style) */
  for (uint64_t i = 0; i < 16 && i < n; ++i) {
    long r, s, e;
    if (fscanf(exp, "%ld %ld %ld", &r, &s, &e) != 3)
      return 2;
    red_b200_outcome o;
    CHECK(red_b200_match(p, text + starts[i], lens[i], RED_B200_VERB_MATCH, RED_B200_STY_INSTANT, &o));
    if (r != o.result || s != (long)o.start || e != (long)o.end) {
      if (bad++ < 5)
        fprintf(stderr, "single %llu: got %d %llu %llu, want %ld %ld %ld\n", (unsigned long long)i, o.result,
                (unsigned long long)o.start, (unsigned long long)o.end, r, s, e);
    }
  }
  /* replace() on the first line */
  {
    long wantLen, wantCount;
    if (fscanf(exp, "%ld %ld", &wantLen, &wantCount) != 2)
      return 2;
    char out[4096];
    size_t outLen = 0;
    uint64_t count = 0;
    CHECK(red_b200_replace(p, text + starts[0], lens[0], "<X>", 3, ~0ull, RED_B200_STY_LAST, out, sizeof out,
                           &outLen, &count));
    if ((long)outLen != wantLen || (long)count != wantCount) {
      fprintf(stderr, "replace: got len %zu count %llu, want %ld %ld\n", outLen, (unsigned long long)count, wantLen,
              wantCount);
      ++bad;
    }
  }
  if (red_b200_program_launch_count(p) <= 0) {
    fprintf(stderr, "no kernel was launched\n");
    ++bad;
  }
  red_b200_program_free(p);
  printf("%s: %llu lines, %u states, %d mismatches\n", bad ? "FAIL" : "OK", (unsigned long long)n, info.state_count,
         bad);
  return bad ? 1 : 0;
}
