/* program.h - host-side view of a serialized RED program and the device image
   built from it.

   The serialized layout is the reference's REDA format
   (reference: include/Serializer.h:42-77, lib/Executable.cpp:159-170):

     bytes 0..3    "REDA"
     4..5, 6..7    major (1), minor (0)
     8..11         FNV-1a/32 of bytes 12..end
     12            format: width in bytes of a table value (1, 2 or 4)
     13            maxChar: largest byte-class number
     14            leaderLen
     16..19        stateCnt
     20..23        initialOff   (BYTE offset of the initial row from the first row)
     24..27        leaderOff    (BYTE offset of the row reached after the leader)
     32..287       equivMap[256]: byte -> class
     288..         leader, zero-padded to a multiple of 8
     then          stateCnt rows of (maxChar + 2) values:
                     [ result | deadEnd << topbit , next[0..maxChar] ]
                   next[] are row offsets in value units.

   The GPU does not walk that layout.  build_image() re-lays it out:

     * states are renumbered  [ rejecting | accepting | pure dead ends ]
       so that "did anything special happen" is one compare of the state
       value against a threshold - no load of the row head per byte;
     * the row head is dropped from the table; results live in a side array
       that is read only when an accepting state is entered;
     * a table entry holds the next state already multiplied into an entry
       offset ("v"), so one step is   v = T[v * unit + class]  ;
     * the class map is stored pre-scaled to byte offsets.
 */

#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace red_b200 {

constexpr size_t kHeaderSize = 288;

struct ParsedProgram {
  uint32_t format = 0;      // 1, 2, 4
  uint32_t maxChar = 0;
  uint32_t classCount = 0;  // maxChar + 1
  uint32_t leaderLen = 0;
  uint32_t stateCount = 0;
  uint32_t initialState = 0;  // row index
  uint32_t leaderState = 0;   // row index
  uint8_t equiv[256];
  std::vector<uint8_t> leader;
  std::vector<int32_t> result;     // per row
  std::vector<uint8_t> deadEnd;    // per row (flag as serialized)
  std::vector<uint8_t> pureDead;   // per row: head == 1 << topbit
  std::vector<uint32_t> next;      // stateCount * classCount, row indices
};

// returns nullptr if fine, else the reference's complaint
const char *check_header(const void *ptr, size_t len);
uint32_t calc_checksum(const void *ptr, size_t len);

// header check + structural validation; false with `why` set on failure
bool parse_program(const void *ptr, size_t len, ParsedProgram &out,
                   std::string &why);

enum TableMode { kTableSmemU16 = 1, kTableGlobalU32 = 2 };

struct DeviceImage {
  TableMode mode = kTableSmemU16;
  uint32_t stateCount = 0;
  uint32_t classCount = 0;
  uint32_t rowStride = 0;   // entries per row (>= classCount)
  uint32_t unitShift = 0;   // entry index of a row = v << unitShift
  uint32_t initV = 0;       // v of the initial state
  uint32_t leaderV = 0;
  uint32_t acceptV = 0;     // acceptV <= v < pureV -> accepting
  uint32_t absV = 0;        // absV <= v < pureV -> accepting and absorbing
  uint32_t pureV = 0;       // v >= pureV -> pure dead end
  uint32_t vPerState = 0;   // v = newIndex * vPerState
  int32_t initResult = 0;
  uint32_t acceptCount = 0;
  uint32_t pureCount = 0;
  int32_t maxResult = 0;
  bool byteRows = false;    // rows indexed by the raw byte (256 columns), identity class map
  bool finalReachable = true;  // a state that can never be left is reachable from the initial one
  std::vector<uint16_t> clsOffset;   // 256: class * entryBytes
  std::vector<uint16_t> table16;     // mode smem
  std::vector<uint32_t> table32;     // mode global
  std::vector<int32_t> results;      // by new index
  std::vector<uint32_t> newToOld;    // new index -> serialized row
  std::vector<uint32_t> oldToNew;    // serialized row -> new index
  std::vector<uint8_t> leader;       // class space
  /* For the long-stream matcher: the states the DFA can be in right behind
     byte b, whatever it was in before - the image under b of all the states
     that can be reached from the initial one - for
     the bytes where those are at most kResetMax (v values, resetCount[b] of
     them; resetCount[b] = 0: more than that).  A chunk's walk can start from
     every one of them behind such a byte and be sure to include the true
     state. */
  static constexpr uint32_t kResetMax = 8;
  std::vector<uint8_t> resetCount;   // [256]
  std::vector<uint32_t> resetStates; // [256 * kResetMax]
  size_t tableBytes() const {
    return mode == kTableSmemU16 ? table16.size() * 2 : table32.size() * 4;
  }
};

// smemBudget: bytes of shared memory the table may take
// visitsByOld (optional, per state of the serialized program): inside each of
// the four groups the states a profiled walk enters most come first, so that
// the leading rows of a global table - the ones the kernels mirror in shared
// memory - are the ones the text actually uses
void build_image(const ParsedProgram &pp, size_t smemBudget, DeviceImage &img,
                 bool allowByteRows = true, const uint64_t *visitsByOld = nullptr);

/* The hot image: what kernels_hot.cu keeps in shared memory.  Up to 255 states
   - the ones a walk over real text spends its time in, by measured visit
   counts where a profile exists, breadth-first from the initial state
   otherwise - get a one-byte CODE and a row indexed by the raw byte, so that a
   step is ONE shared-memory load:   code = T[code * kHotStride + byte].
   Codes follow the order of the renumbered states, so the thresholds of the
   full image carry over (code >= cAccept: accepting, ...); every transition
   that leaves the hot set leads to the code nHot ("sink", a row of its own
   that loops to itself), and the lane then continues on the full table.
   A row is 256 entries + 4 bytes (kHotStrideWide), or - for text that is
   ASCII, which the profile tells - 128 entries + 4 bytes (kHotStrideAscii: a
   byte >= 128 then counts as leaving the hot set, and half the shared memory
   is free for text).  The 4 bytes make consecutive rows start one bank apart,
   which spreads the lanes of a warp that sit in different rows over the banks
   (simulated on the bench corpus: 1.5 wavefronts per lookup, 2.0 without). */
constexpr uint32_t kHotStrideWide = 260;
constexpr uint32_t kHotStrideAscii = 132;
constexpr uint32_t kHotMax = 255;
/* The long-stream matcher's image may hold more states than a byte can name:
   two-byte entries then (rows of 2 * columns + 4 bytes), as many states as the
   kernel's shared memory takes beside its text stages. */
constexpr uint32_t kHotMaxWide = 1024;

struct HotImage {
  uint32_t nHot = 0;              // codes 0 .. nHot-1; nHot = sink
  uint32_t cAccept = 0;           // codes >= cAccept (and < cPure): accepting
  uint32_t cAbs = 0;              // codes >= cAbs (and < cPure): accepting, never left
  uint32_t cPure = 0;             // codes >= cPure (and < nHot): pure dead ends
  uint32_t initCode = 0;
  uint32_t stride = kHotStrideWide; // bytes per row
  uint32_t coverage = 1000;       // per mille of the counted state entries that the hot set holds
  bool acceptSticks = false;      // rows of accepting codes loop to themselves (see build_hot_image)
  uint32_t entryBytes = 1;        // 1: one-byte codes (at most kHotMax states); 2: two-byte codes
  std::vector<uint8_t> table;     // (nHot + 1) * stride, padded to 16 bytes
  std::vector<uint32_t> codeToV;  // max(256, nHot + 1) entries: v of the full image
  std::vector<uint8_t> codeOf;    // by renumbered state index (entryBytes each); nHot = not hot
};

/* visits: per renumbered state index, or nullptr (breadth-first order then).
   acceptSticks: the image of style Instant, whose walk ends in the first
   accepting state - every accepting code's row loops to itself, so that a code
   >= cAccept, once reached, is still there at the end of a 16-byte chunk and
   the fast pass needs no running maximum (the sink and the absorbing and dead
   states stick anyway). */
void build_hot_image(const ParsedProgram &pp, const DeviceImage &img,
                     const uint64_t *visits, uint32_t maxHot, bool asciiRows, bool acceptSticks,
                     HotImage &hot, bool wideCodes = false);

} // namespace red_b200
