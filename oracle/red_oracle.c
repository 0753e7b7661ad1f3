/* red_oracle.c - CPU restatement of RED's matcher over the serialized DFA.

   TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
   bench.py's cpu_baseline / --impl reference legs may load this.  The
   product (one_b200/) never links, loads or calls anything in oracle/.

   What is restated (reference = /root/reference/quol/red):
     - serialized program layout     include/Serializer.h:42-77
     - header validation + checksum  lib/Serializer.cpp:270-306,
                                     include/Fnv.h:45-63 (FNV-1a, 32 bit)
     - where the state rows start    lib/Executable.cpp:159-170
     - state access                  include/Proxy.h:117-152
     - leader helpers                include/Matcher.h:333-360
     - check / match / scan / search include/Matcher.h:363-641
     - replace / matchAll            include/Matcher.h:643-767
     - StatefulMatcher               include/Matcher.h:770-794,
                                     lib/Matcher.cpp:106-158
     - Red::collect                  lib/Red.cpp:108-122

   Parity pin: tests/test_oracle_golden.py checks this file against
   (1) the expected values transcribed from the reference's own unit tests
       (test/matcher.cpp, test/red.cpp, test/omnibus.cpp) and
   (2) tests/golden/*.json, outputs of the reference itself (oracle/_ref,
       the reference compiled unmodified) produced by tools/gen_golden.py.

   The program is executed exactly as serialized: a state is a byte offset
   from `base`; row = [resultAndDeadEnd, next[0..maxChar]] of 1/2/4-byte
   values; next[] holds offsets in units of the value size.
 */

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { STY_INSTANT = 1, STY_FIRST = 2, STY_TANGENT = 3, STY_LAST = 4,
       STY_FULL = 5 };
enum { VERB_CHECK = 0, VERB_MATCH = 1, VERB_SCAN = 2, VERB_SEARCH = 3 };

#define HDR_SIZE 288u /* sizeof(FileHeader): 32 fixed bytes + equivMap[256] */

typedef struct {
  uint8_t *buf;         /* private copy of the serialized program */
  uint64_t len;
  const uint8_t *equiv; /* buf + 32 */
  const uint8_t *leader;
  const uint8_t *base;  /* first state row */
  uint32_t fmt;         /* 1, 2 or 4 */
  uint32_t maxChar;
  uint32_t leaderLen;
  uint32_t stateCnt;
  uint32_t initialOff;
  uint32_t leaderOff;
} red_oracle_prog;

typedef struct {
  int32_t result;
  uint64_t start;
  uint64_t end;
} red_oracle_outcome;

static uint32_t rd32(const uint8_t *p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}

static uint16_t rd16(const uint8_t *p) {
  uint16_t v;
  memcpy(&v, p, 2);
  return v;
}

/* FNV-1a over everything from the format byte (offset 12) to the end */
uint32_t red_oracle_checksum(const void *ptr, uint64_t len) {
  const uint8_t *p = (const uint8_t *)ptr;
  uint32_t h = 0x811c9dc5u;
  for (uint64_t ii = 12; ii < len; ++ii) {
    h ^= p[ii];
    h *= 0x01000193u;
  }
  return h;
}

/* returns NULL if good, else the complaint (same wording as the reference) */
const char *red_oracle_check_header(const void *ptr, uint64_t len) {
  const uint8_t *p = (const uint8_t *)ptr;
  if (len < HDR_SIZE)
    return "Serialized DFA: header too short";
  if (memcmp(p, "REDA", 4) != 0)
    return "Serialized DFA: bad magic number";
  if ((rd16(p + 4) != 1) || (rd16(p + 6) != 0))
    return "Serialized DFA: unrecognized version";
  uint32_t want = rd32(p + 8);
  uint32_t got = red_oracle_checksum(ptr, len);
  if (want != got) {
    if (want == __builtin_bswap32(got))
      return "serialized DFA: foreign endian-ness";
    return "serialized DFA: checksum mismatch";
  }
  uint8_t fmt = p[12];
  if ((fmt != 1) && (fmt != 2) && (fmt != 4))
    return "Serialized DFA: unsupported format";
  return NULL;
}

red_oracle_prog *red_oracle_load(const void *ptr, uint64_t len,
                                 const char **why) {
  const char *msg = ptr ? red_oracle_check_header(ptr, len)
                        : "serialized dfa is empty";
  if (msg) {
    if (why) *why = msg;
    return NULL;
  }
  red_oracle_prog *pg = (red_oracle_prog *)calloc(1, sizeof(*pg));
  pg->buf = (uint8_t *)malloc(len);
  memcpy(pg->buf, ptr, len);
  pg->len = len;
  const uint8_t *p = pg->buf;
  pg->fmt = p[12];
  pg->maxChar = p[13];
  pg->leaderLen = p[14];
  pg->stateCnt = rd32(p + 16);
  pg->initialOff = rd32(p + 20);
  pg->leaderOff = rd32(p + 24);
  pg->equiv = p + 32;
  pg->leader = p + HDR_SIZE;
  pg->base = p + HDR_SIZE + ((pg->leaderLen + 7u) & ~7u);
  return pg;
}

void red_oracle_free(red_oracle_prog *pg) {
  if (pg) {
    free(pg->buf);
    free(pg);
  }
}

uint32_t red_oracle_info(const red_oracle_prog *pg, int what) {
  switch (what) {
  case 0: return pg->fmt;
  case 1: return pg->maxChar;
  case 2: return pg->leaderLen;
  case 3: return pg->stateCnt;
  case 4: return pg->initialOff;
  case 5: return pg->leaderOff;
  default: return 0;
  }
}

/* does text[pos..] begin with the leader (compared in class space)? */
static int leader_at(const red_oracle_prog *pg, const uint8_t *t, uint64_t pos,
                     uint64_t len, uint64_t *matched) {
  uint64_t ii = 0;
  int ok = 1;
  for (; ii < pg->leaderLen; ++ii) {
    if ((pos + ii >= len) || (pg->leader[ii] != pg->equiv[t[pos + ii]])) {
      ok = 0;
      break;
    }
  }
  if (matched) *matched = ii;
  return ok;
}

/* One set of walkers per value width.  S is the current state as a byte
   offset from base; word 0 of the row is result|deadEnd<<topbit. */
#define DEFINE_WALKERS(SFX, T, TOPBIT)                                        \
  static inline T rowhead_##SFX(const red_oracle_prog *pg, uint64_t s) {      \
    T v;                                                                      \
    memcpy(&v, pg->base + s, sizeof(T));                                      \
    return v;                                                                 \
  }                                                                           \
  static inline uint64_t step_##SFX(const red_oracle_prog *pg, uint64_t s,    \
                                    uint8_t cls) {                            \
    T v;                                                                      \
    memcpy(&v, pg->base + s + sizeof(T) * (1u + cls), sizeof(T));             \
    return (uint64_t)v * sizeof(T);                                           \
  }                                                                           \
  static inline int32_t res_##SFX(T head) {                                   \
    return (int32_t)(head & (T)~((T)1 << TOPBIT));                            \
  }                                                                           \
  static inline int pure_dead_##SFX(T head) {                                 \
    return head == (T)((T)1 << TOPBIT);                                       \
  }                                                                           \
                                                                              \
  /* anchored walk from state s0 over t[from..len); returns result and, if    \
     oc != NULL, positions relative to t[0].  Shared by all four verbs. */    \
  static int32_t walk_##SFX(const red_oracle_prog *pg, const uint8_t *t,      \
                            uint64_t from, uint64_t len, uint64_t s0,         \
                            int style, red_oracle_outcome *oc) {              \
    const uint64_t init = pg->initialOff;                                     \
    uint64_t s = s0;                                                          \
    int32_t result = res_##SFX(rowhead_##SFX(pg, s));                         \
    int32_t prev = 0;                                                         \
    uint64_t mstart = from, mend = from;                                      \
    for (uint64_t idx = from; idx < len; ++idx) {                             \
      uint8_t cls = pg->equiv[t[idx]];                                        \
      uint64_t nx = step_##SFX(pg, s, cls);                                   \
      if ((s == init) && (nx != s))                                           \
        mstart = idx;                                                         \
      s = nx;                                                                 \
      T head = rowhead_##SFX(pg, s);                                          \
      result = res_##SFX(head);                                               \
      if (result > 0) {                                                       \
        if (style == STY_FIRST) {                                             \
          if (prev && (result != prev)) {                                     \
            result = prev;                                                    \
            break;                                                            \
          }                                                                   \
          prev = result;                                                      \
        }                                                                     \
        mend = idx + 1;                                                       \
        if (style == STY_INSTANT)                                             \
          break;                                                              \
        if ((style == STY_TANGENT) || (style == STY_LAST))                    \
          prev = result;                                                      \
      }                                                                       \
      else {                                                                  \
        if ((style == STY_FIRST) && (prev > 0)) {                             \
          result = prev;                                                      \
          break;                                                              \
        }                                                                     \
        if ((style == STY_TANGENT) && (prev > 0))                             \
          break;                                                              \
        if (pure_dead_##SFX(head))                                            \
          break;                                                              \
      }                                                                       \
    }                                                                         \
    if ((style == STY_TANGENT) || (style == STY_LAST))                        \
      if ((result == 0) && (prev > 0))                                        \
        result = prev;                                                        \
    if (oc) {                                                                 \
      oc->result = result;                                                    \
      oc->start = result ? mstart : 0;                                        \
      oc->end = result ? mend : 0;                                            \
    }                                                                         \
    return result;                                                            \
  }

DEFINE_WALKERS(1, uint8_t, 7)
DEFINE_WALKERS(2, uint16_t, 15)
DEFINE_WALKERS(4, uint32_t, 31)

static int32_t walk(const red_oracle_prog *pg, const uint8_t *t, uint64_t from,
                    uint64_t len, uint64_t s0, int style,
                    red_oracle_outcome *oc) {
  switch (pg->fmt) {
  case 1: return walk_1(pg, t, from, len, s0, style, oc);
  case 2: return walk_2(pg, t, from, len, s0, style, oc);
  default: return walk_4(pg, t, from, len, s0, style, oc);
  }
}

static void set_fail(red_oracle_outcome *oc) {
  oc->result = 0;
  oc->start = 0;
  oc->end = 0;
}

/* check: Matcher.h:363-410.  With the leader optimisation the walk starts
   after the leader, from the state the leader leads to. */
int32_t red_oracle_check(const red_oracle_prog *pg, const void *text,
                         uint64_t len, int style, int do_leader) {
  const uint8_t *t = (const uint8_t *)text;
  if (do_leader) {
    if (!leader_at(pg, t, 0, len, NULL))
      return 0;
    return walk(pg, t, pg->leaderLen, len, pg->leaderOff, style, NULL);
  }
  return walk(pg, t, 0, len, pg->initialOff, style, NULL);
}

/* match: Matcher.h:413-495.  The leader only gates; the walk still starts at
   the initial state so that start_ is where the initial state is left. */
void red_oracle_match(const red_oracle_prog *pg, const void *text,
                      uint64_t len, int style, int do_leader,
                      red_oracle_outcome *oc) {
  const uint8_t *t = (const uint8_t *)text;
  if (do_leader && !leader_at(pg, t, 0, len, NULL)) {
    set_fail(oc);
    return;
  }
  walk(pg, t, 0, len, pg->initialOff, style, oc);
}

/* scan: Matcher.h:498-554.  Quirk kept on purpose: with the leader on, a
   failed leader comparison leaves the cursor where the comparison stopped
   (compareThrough takes the iterator by reference), and a successful one
   leaves it after the leader, so candidate start positions are skipped. */
int32_t red_oracle_scan(const red_oracle_prog *pg, const void *text,
                        uint64_t len, int style, int do_leader) {
  const uint8_t *t = (const uint8_t *)text;
  red_oracle_outcome tmp;
  walk(pg, t, 0, 0, pg->initialOff, STY_FULL, &tmp); /* initial result */
  int32_t result = tmp.result;
  for (uint64_t pos = 0; pos < len; ++pos) {
    uint64_t s0 = pg->initialOff;
    if (do_leader) {
      uint64_t adv;
      int ok = leader_at(pg, t, pos, len, &adv);
      pos += adv;
      if (!ok)
        continue;
      s0 = pg->leaderOff;
    }
    /* the inner loop of scanCore is the check walk without the final
       "styLast falls back to prev" being skipped: it is applied too */
    result = walk(pg, t, pos, len, s0, style, NULL);
    if (result > 0)
      return result;
  }
  return result;
}

/* search: Matcher.h:557-641 */
void red_oracle_search(const red_oracle_prog *pg, const void *text,
                       uint64_t len, int style, int do_leader,
                       red_oracle_outcome *oc) {
  const uint8_t *t = (const uint8_t *)text;
  set_fail(oc);
  for (uint64_t pos = 0; pos < len; ++pos) {
    if (do_leader && !leader_at(pg, t, pos, len, NULL))
      continue;
    red_oracle_outcome cur;
    walk(pg, t, pos, len, pg->initialOff, style, &cur);
    if (cur.result > 0) {
      *oc = cur;
      return;
    }
  }
  /* empty text, or nothing found: the reference reports the initial state's
     result for an empty input (the loop body never runs) */
  if (len == 0) {
    red_oracle_outcome cur;
    walk(pg, t, 0, 0, pg->initialOff, style, &cur);
    *oc = cur;
  }
}

/* one call per line, as a RED user would loop; lines = text[starts[i] ..
   starts[i]+lens[i]).  starts_out / ends_out may be NULL. */
void red_oracle_batch(const red_oracle_prog *pg, const void *text,
                      const uint64_t *starts, const uint32_t *lens, uint64_t n,
                      int verb, int style, int do_leader, int32_t *results,
                      uint32_t *starts_out, uint32_t *ends_out) {
  const uint8_t *t = (const uint8_t *)text;
  for (uint64_t ii = 0; ii < n; ++ii) {
    red_oracle_outcome oc;
    set_fail(&oc);
    const uint8_t *line = t + starts[ii];
    switch (verb) {
    case VERB_CHECK:
      oc.result = red_oracle_check(pg, line, lens[ii], style, do_leader);
      break;
    case VERB_SCAN:
      oc.result = red_oracle_scan(pg, line, lens[ii], style, do_leader);
      break;
    case VERB_MATCH:
      red_oracle_match(pg, line, lens[ii], style, do_leader, &oc);
      break;
    default:
      red_oracle_search(pg, line, lens[ii], style, do_leader, &oc);
      break;
    }
    results[ii] = oc.result;
    if (starts_out) starts_out[ii] = (uint32_t)oc.start;
    if (ends_out) ends_out[ii] = (uint32_t)oc.end;
  }
}

/* Red::collect(): repeated search<styLast,false> from just after the
   previous match (lib/Red.cpp:108-122).  Returns the number found; stores at
   most cap of them. */
uint64_t red_oracle_collect(const red_oracle_prog *pg, const void *text,
                            uint64_t len, uint64_t cap, int32_t *results,
                            uint64_t *starts, uint64_t *ends) {
  const uint8_t *t = (const uint8_t *)text;
  uint64_t cnt = 0;
  uint64_t off = 0;
  while (off < len) {
    red_oracle_outcome oc;
    red_oracle_search(pg, t + off, len - off, STY_LAST, 0, &oc);
    if (oc.result <= 0)
      break;
    if (cnt < cap) {
      results[cnt] = oc.result;
      starts[cnt] = off + oc.start;
      ends[cnt] = off + oc.end;
    }
    ++cnt;
    off += oc.end;
  }
  return cnt;
}

/* matchAll(): one anchored walk that reports every run of positions with the
   same positive result (Matcher.h:711-767; called with the leader on, which
   only gates - a text that does not begin with the leader yields nothing).
   Returns the number found; stores at most cap of them. */
#define DEFINE_MATCH_ALL(SFX, T)                                              \
  static uint64_t match_all_##SFX(const red_oracle_prog *pg, const uint8_t *t,\
                                  uint64_t len, uint64_t cap, int32_t *results,\
                                  uint64_t *starts, uint64_t *ends) {         \
    const uint64_t init = pg->initialOff;                                     \
    uint64_t s = init, cnt = 0, mstart = 0;                                   \
    int32_t prev = 0;                                                         \
    for (uint64_t idx = 0; idx < len; ++idx) {                                \
      uint64_t nx = step_##SFX(pg, s, pg->equiv[t[idx]]);                     \
      if ((s == init) && (nx != s))                                           \
        mstart = idx;                                                         \
      s = nx;                                                                 \
      T head = rowhead_##SFX(pg, s);                                          \
      int32_t result = res_##SFX(head);                                       \
      if (result > 0) {                                                       \
        if (result == prev) {                                                 \
          if (cnt - 1 < cap)                                                  \
            ends[cnt - 1] = idx + 1;                                          \
        }                                                                     \
        else {                                                                \
          prev = result;                                                      \
          if (cnt < cap) {                                                    \
            results[cnt] = result;                                            \
            starts[cnt] = mstart;                                             \
            ends[cnt] = idx + 1;                                              \
          }                                                                   \
          ++cnt;                                                              \
        }                                                                     \
      }                                                                       \
      else {                                                                  \
        if (pure_dead_##SFX(head))                                            \
          break;                                                              \
        prev = 0;                                                             \
      }                                                                       \
    }                                                                         \
    return cnt;                                                               \
  }

DEFINE_MATCH_ALL(1, uint8_t)
DEFINE_MATCH_ALL(2, uint16_t)
DEFINE_MATCH_ALL(4, uint32_t)

uint64_t red_oracle_match_all(const red_oracle_prog *pg, const void *text,
                              uint64_t len, uint64_t cap, int32_t *results,
                              uint64_t *starts, uint64_t *ends) {
  const uint8_t *t = (const uint8_t *)text;
  if (!leader_at(pg, t, 0, len, NULL))
    return 0;
  switch (pg->fmt) {
  case 1: return match_all_1(pg, t, len, cap, results, starts, ends);
  case 2: return match_all_2(pg, t, len, cap, results, starts, ends);
  default: return match_all_4(pg, t, len, cap, results, starts, ends);
  }
}

/* replace(): rewrite the text, substituting `repl` for each non-overlapping
   match found by trying every position in order (replaceCore,
   Matcher.h:643-707; the runtime-style entry points use the leader,
   lib/Matcher.cpp:72-92).  Unlike matchCore the inner walk keeps no start
   position, Full forgets a match as soon as a later state rejects, and a text
   position is only ever consumed as part of a match or copied through.
   Writes at most cap bytes to out; *out_len is the full length.  Returns the
   number of replacements. */
#define DEFINE_REPLACE(SFX, T)                                                \
  static uint64_t replace_##SFX(const red_oracle_prog *pg, const uint8_t *t,  \
                                uint64_t len, const uint8_t *repl,            \
                                uint64_t repl_len, uint64_t max, int style,   \
                                uint8_t *out, uint64_t cap,                   \
                                uint64_t *out_len) {                          \
    uint64_t cnt = 0, o = 0, pos = 0;                                         \
    while (pos < len) {                                                       \
      if (cnt >= max) {                                                       \
        for (; pos < len; ++pos, ++o)                                         \
          if (o < cap) out[o] = t[pos];                                       \
        break;                                                                \
      }                                                                       \
      int have = 0;                                                           \
      uint64_t found = 0;                                                     \
      if (leader_at(pg, t, pos, len, NULL)) {                                 \
        uint64_t s = pg->initialOff;                                          \
        int32_t prev = 0;                                                     \
        for (uint64_t idx = pos; idx < len; ++idx) {                          \
          s = step_##SFX(pg, s, pg->equiv[t[idx]]);                           \
          T head = rowhead_##SFX(pg, s);                                      \
          int32_t result = res_##SFX(head);                                   \
          if (result > 0) {                                                   \
            if (style == STY_FIRST) {                                         \
              if (prev && (result != prev))                                   \
                break;                                                        \
              prev = result;                                                  \
            }                                                                 \
            have = 1;                                                         \
            found = idx;                                                      \
            if (style == STY_INSTANT)                                         \
              break;                                                          \
          }                                                                   \
          else {                                                              \
            if (style == STY_FULL)                                            \
              have = 0;                                                       \
            if ((((style == STY_FIRST) || (style == STY_TANGENT)) && have) || \
                pure_dead_##SFX(head))                                        \
              break;                                                          \
          }                                                                   \
        }                                                                     \
      }                                                                       \
      if (have) {                                                             \
        for (uint64_t ii = 0; ii < repl_len; ++ii, ++o)                       \
          if (o < cap) out[o] = repl[ii];                                     \
        pos = found + 1;                                                      \
        ++cnt;                                                                \
      }                                                                       \
      else {                                                                  \
        if (o < cap) out[o] = t[pos];                                         \
        ++o;                                                                  \
        ++pos;                                                                \
      }                                                                       \
    }                                                                         \
    *out_len = o;                                                             \
    return cnt;                                                               \
  }

DEFINE_REPLACE(1, uint8_t)
DEFINE_REPLACE(2, uint16_t)
DEFINE_REPLACE(4, uint32_t)

uint64_t red_oracle_replace(const red_oracle_prog *pg, const void *text,
                            uint64_t len, const void *repl, uint64_t repl_len,
                            uint64_t max, int style, void *out, uint64_t cap,
                            uint64_t *out_len) {
  const uint8_t *t = (const uint8_t *)text;
  const uint8_t *r = (const uint8_t *)repl;
  switch (pg->fmt) {
  case 1: return replace_1(pg, t, len, r, repl_len, max, style, (uint8_t *)out, cap, out_len);
  case 2: return replace_2(pg, t, len, r, repl_len, max, style, (uint8_t *)out, cap, out_len);
  default: return replace_4(pg, t, len, r, repl_len, max, style, (uint8_t *)out, cap, out_len);
  }
}

/* StatefulMatcher (Matcher.h:770-794): the state is carried by the caller.
   *state is a byte offset from base; pass RED_ORACLE_FRESH to start from the
   initial state (the constructor, lib/Matcher.cpp:106-139).  Feeds n bytes;
   per_byte[i] (if not NULL) is what advance() returned for byte i.  Returns
   result() after the last byte - for n == 0 the result of the state itself. */
#define RED_ORACLE_FRESH (~(uint64_t)0)

#define DEFINE_ADVANCE(SFX, T)                                                \
  static int32_t advance_##SFX(const red_oracle_prog *pg, uint64_t *state,    \
                               const uint8_t *t, uint64_t n,                  \
                               int32_t *per_byte) {                           \
    uint64_t s = (*state == RED_ORACLE_FRESH) ? pg->initialOff : *state;      \
    int32_t result = res_##SFX(rowhead_##SFX(pg, s));                         \
    for (uint64_t ii = 0; ii < n; ++ii) {                                     \
      s = step_##SFX(pg, s, pg->equiv[t[ii]]);                                \
      result = res_##SFX(rowhead_##SFX(pg, s));                               \
      if (per_byte) per_byte[ii] = result;                                    \
    }                                                                         \
    *state = s;                                                               \
    return result;                                                            \
  }

DEFINE_ADVANCE(1, uint8_t)
DEFINE_ADVANCE(2, uint16_t)
DEFINE_ADVANCE(4, uint32_t)

int32_t red_oracle_advance(const red_oracle_prog *pg, uint64_t *state,
                           const void *bytes, uint64_t n, int32_t *per_byte) {
  const uint8_t *t = (const uint8_t *)bytes;
  switch (pg->fmt) {
  case 1: return advance_1(pg, state, t, n, per_byte);
  case 2: return advance_2(pg, state, t, n, per_byte);
  default: return advance_4(pg, state, t, n, per_byte);
  }
}
