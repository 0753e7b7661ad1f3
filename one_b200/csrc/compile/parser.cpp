/* parser.cpp - RED's pattern languages -> NFA.

   Tokenizer and recursive-descent parser for the regex dialect of the
   reference (grammar and escapes: reference doc/Usage.md "Regular Expression
   Syntax"; behaviour: lib/Scanner.cpp, lib/Parser.cpp), plus globs and exact
   strings.  Accept/reject decisions and the automaton built are kept
   identical to the reference; error texts are ours. */

#include "parser.h"

#include <limits>

#include "../../../include/red_b200.h"

namespace red_b200 {
namespace rx {

namespace {

enum TokKind { kEnd, kFlags, kBytes, kRepeat, kBar, kOpen, kClose };

struct Tok {
  TokKind kind = kEnd;
  size_t pos = 0;
  int min = 0, max = 0;  // kRepeat (max < 0: unbounded)
  uint32_t flags = 0;    // kFlags
  ByteSet set;           // kBytes
};

int hexval(uint8_t c) {
  if ((c >= '0') && (c <= '9')) return c - '0';
  if ((c >= 'A') && (c <= 'F')) return c - 'A' + 10;
  if ((c >= 'a') && (c <= 'f')) return c - 'a' + 10;
  return -1;
}

// \d \D \s \S \w \W expand to sets; any other escaped byte stands for itself
void addClassOrByte(ByteSet &set, int c, bool escaped) {
  if (!escaped) {
    set.add(c);
    return;
  }
  ByteSet inv;
  switch (c) {
  case 'd': set.addRange('0', '9'); break;
  case 'D': inv.addRange('0', '9'); inv.invert(); set.merge(inv); break;
  case 's':
    for (int w : {'\t', '\n', '\v', '\f', '\r', ' '}) set.add(w);
    break;
  case 'S':
    for (int w : {'\t', '\n', '\v', '\f', '\r', ' '}) inv.add(w);
    inv.invert(); set.merge(inv);
    break;
  case 'w':
    set.addRange('0', '9'); set.addRange('A', 'Z'); set.addRange('a', 'z'); set.add('_');
    break;
  case 'W':
    inv.addRange('0', '9'); inv.addRange('A', 'Z'); inv.addRange('a', 'z'); inv.add('_');
    inv.invert(); set.merge(inv);
    break;
  default: set.add(c); break;
  }
}

class Lexer {
public:
  void reset(const uint8_t *p, size_t n) { beg_ = cur_ = p; end_ = p + n; count_ = 0; }
  size_t tokens() const { return count_; }

  Tok next() {
    ++count_;
    Tok t;
    if (cur_ >= end_) {
      t.kind = kEnd; t.pos = pos();
      return t;
    }
    switch (*cur_) {
    case '(': ++cur_; t.kind = kOpen; t.pos = pos(); return t;
    case ')': ++cur_; t.kind = kClose; t.pos = pos(); return t;
    case '|': ++cur_; t.kind = kBar; t.pos = pos(); return t;
    case '*': ++cur_; return rep(0, -1);
    case '+': ++cur_; return rep(1, -1);
    case '?': ++cur_; return rep(0, 1);
    case '.': ++cur_; t.kind = kBytes; t.pos = pos(); t.set.fill(); return t;
    case '[': ++cur_; return bracket();
    case '{': ++cur_; return braces();
    default: return ordinary();
    }
  }

private:
  size_t pos() const { return static_cast<size_t>(cur_ - beg_); }

  Tok rep(int mn, int mx) {
    Tok t; t.kind = kRepeat; t.pos = pos(); t.min = mn; t.max = mx;
    return t;
  }

  // one (possibly escaped) byte; -1 at end of input.  `escaped` is set for
  // backslash sequences that are not a literal byte (\d, \i, \1 ...)
  int oneByte(bool &escaped) {
    if (cur_ >= end_)
      return -1;
    uint8_t c = *cur_++;
    escaped = false;
    if (c != '\\')
      return c;
    if (cur_ >= end_)
      fail_parse("partial backslash escape", pos());
    c = *cur_++;
    switch (c) {
    case '0': return '\0';
    case 'a': return '\a';
    case 'b': return '\b';
    case 'n': return '\n';
    case 'r': return '\r';
    case 't': return '\t';
    case 'v': return '\v';
    case '\\': return '\\';
    case 'x': {
      if (cur_ + 2 > end_)
        fail_parse("unterminated hex escape", pos());
      int hi = hexval(cur_[0]), lo = hexval(cur_[1]);
      if ((hi < 0) || (lo < 0))
        fail_parse("invalid hex escape", pos());
      cur_ += 2;
      return (hi << 4) | lo;
    }
    default:
      escaped = true;
      return c;
    }
  }

  // [...] : same walk as the reference's scanSet (lib/Scanner.cpp:137), whose
  // corner cases (leading ']' or '-', "a-]" , escaped class as a range end)
  // are part of the dialect
  Tok bracket() {
    Tok t; t.kind = kBytes; t.pos = pos();
    bool xe = false, ye = false, invert = false;
    int x = oneByte(xe);
    if (x < 0) fail_parse("incomplete character class", pos());
    if (!xe && (x == '^')) {
      invert = true;
      x = oneByte(xe);
      if (x < 0) fail_parse("incomplete inverted character class", pos());
    }
    if (!xe && ((x == '-') || (x == ']'))) {
      t.set.add(x);
      x = oneByte(xe);
      if (x < 0) fail_parse("unfinished character class", pos());
    }
    while (xe || (x != ']')) {
      int y = oneByte(ye);
      if (y < 0) fail_parse("unexpected end of character class", pos());
      if (!ye && (y == '-')) {
        y = oneByte(ye);
        if (y < 0) fail_parse("unfinished character class range", pos());
        if (!ye && (y == ']')) {
          t.set.add(x);
          t.set.add('-');
          x = y; xe = ye;
          continue;
        }
        if (x > y) fail_parse("backward character class range", pos());
        t.set.addRange(x, y);
        x = oneByte(xe);
        if (x < 0) fail_parse("incomplete character class range", pos());
      }
      else {
        addClassOrByte(t.set, x, xe);
        x = y; xe = ye;
      }
    }
    if (invert)
      t.set.invert();
    return t;
  }

  // {n} {n,m} {n,} {,m} {,} : reference scanClosure (lib/Scanner.cpp:196)
  Tok braces() {
    Tok t; t.kind = kRepeat; t.pos = pos();
    t.min = -1; t.max = 0;
    bool haveMax = false;
    uint8_t c = 0;
    while (cur_ < end_) {
      c = *cur_++;
      if (c == '}')
        break;
      if (c == ',') {
        if (t.min >= 0) fail_parse("extra comma in count", pos());
        t.min = t.max; t.max = 0; haveMax = false;
      }
      else if ((c >= '0') && (c <= '9')) {
        t.max = 10 * t.max + (c - '0');
        haveMax = true;
      }
      else
        fail_parse("non-digit in count", pos());
    }
    if (c != '}') fail_parse("unclosed brace count", pos());
    if ((t.min < 0) && (t.max == 0)) fail_parse("illegal count", pos());
    if (!haveMax) t.max = -1;
    if ((t.min == 0) && (t.max == 0)) fail_parse("zero brace count", pos());
    if ((t.max >= 0) && (t.max < t.min)) fail_parse("backwards brace count", pos());
    if (t.min < 0) t.min = t.max;
    return t;
  }

  Tok ordinary() {
    bool esc = false;
    int c = oneByte(esc);
    Tok t; t.pos = pos();
    if (c < 0) {
      t.kind = kEnd;
      return t;
    }
    if (esc) {
      if ((c >= '1') && (c <= '9'))
        fail_parse("dfa does not support backreferences", pos());
      if (c == 'i') {
        t.kind = kFlags; t.flags = RED_B200_F_IGNORE_CASE;
        return t;
      }
    }
    t.kind = kBytes;
    addClassOrByte(t.set, c, esc);
    return t;
  }

  const uint8_t *beg_ = nullptr, *cur_ = nullptr, *end_ = nullptr;
  size_t count_ = 0;
};

} // anonymous

struct PatternParser::Impl {
  Nfa nfa;
  Lexer lex;
  Tok tok;
  uint32_t flags = 0;
  int depth = 0;
  bool begun = false;
  uint64_t parenBudget;

  Impl(uint64_t states, uint64_t parens) : nfa(states), parenBudget(parens) {}

  void advance() { tok = lex.next(); }

  // expr := part ('|' part)*
  NodeId expr() {
    NodeId s = part();
    if (!s) return 0;
    while (tok.kind == kBar) {
      advance();
      NodeId o = part();
      if (!o) return 0;
      s = nfa.alternate(s, o);
    }
    return s;
  }

  // part := multi multi*
  NodeId part() {
    begun = false;
    NodeId s = multi();
    if (!s) return 0;
    while ((tok.kind == kBytes) || (tok.kind == kOpen)) {
      NodeId m = multi();
      if (!m) return 0;
      s = nfa.sequence(s, m);
    }
    return s;
  }

  // multi := unit (repeat | flags)? flags*
  NodeId multi() {
    NodeId s = unit();
    if (!s) return 0;
    if (tok.kind == kRepeat)
      s = nfa.repeat(s, tok.min, tok.max);
    else if (tok.kind == kFlags)
      flags |= tok.flags;
    else
      return s;
    advance();
    while (tok.kind == kFlags) {
      flags |= tok.flags;
      advance();
    }
    return s;
  }

  // unit := flags* (bytes | '(' expr ')' | <empty>)
  NodeId unit() {
    while (tok.kind == kFlags) {
      flags |= tok.flags;
      advance();
    }
    NodeId s = 0;
    switch (tok.kind) {
    case kEnd:
      if (begun) fail_parse("end reached, expected chars or (", tok.pos);
      return nfa.goalNode();  // empty matches empty
    case kBytes:
      begun = true;
      s = nfa.oneOf(tok.set);
      break;
    case kBar:
      if (begun) fail_parse("expected chars or (, got |", tok.pos);
      return nfa.goalNode();
    case kOpen:
      advance();
      if (parenBudget == 0) fail_limit("nesting budget exceeded");
      --parenBudget;
      ++depth;
      s = expr();
      --depth;
      ++parenBudget;
      if (tok.kind != kClose) fail_parse("close parenthesis not found", tok.pos);
      begun = true;
      break;
    case kClose:
      if (begun || (depth <= 0)) fail_parse("unexpected )", tok.pos);
      return nfa.goalNode();
    default:
      fail_parse("expected chars or (", tok.pos);
    }
    advance();
    return s;
  }

  static void checkResult(int32_t result) {
    if (result <= 0) fail_api("result must be positive");
    if (result > std::numeric_limits<int32_t>::max() - 256) fail_api("result too large");
  }

  // wrap a parsed fragment: case folding, loose ends, end mark; then join it
  // to the automaton's initial node (reference: Parser::add, lib/Parser.cpp:68-83)
  void finishPattern(NodeId s, NodeId startWild, int32_t result, bool endMarkDone) {
    if (flags & RED_B200_F_IGNORE_CASE)
      s = nfa.foldCase(s);
    if (startWild)
      s = nfa.sequence(startWild, s);
    if (flags & RED_B200_F_LOOSE_END) {
      s = nfa.sequence(s, nfa.wildcard());
      if (endMarkDone)
        s = nfa.sequence(s, nfa.endMark(result));
    }
    if (!endMarkDone)
      s = nfa.sequence(s, nfa.endMark(result));
    nfa.joinInitial(s);
  }

  void addRegex(const uint8_t *p, size_t n, int32_t result, uint32_t f) {
    checkResult(result);
    nfa.setGoal(result);
    flags = f;
    depth = 0;
    NodeId startWild = (flags & RED_B200_F_LOOSE_START) ? nfa.wildcard() : 0;
    lex.reset(p, n);
    advance();
    NodeId s = expr();
    if (!s) fail_parse("uncategorized parsing error");
    if (tok.kind != kEnd) fail_parse("trailing unparsed input");
    finishPattern(s, startWild, result, false);
  }

  // glob character class; simplified sibling of bracket()
  // (reference: Parser::parseClass, lib/Parser.cpp:398)
  NodeId globClass(const uint8_t *&ptr, const uint8_t *beg, const uint8_t *end) {
    bool invert = false;
    ByteSet set;
    auto need = [&](const char *what) {
      if (++ptr >= end) fail_parse(what, ptr - beg);
      return *ptr;
    };
    uint8_t x = need("unclosed glob class");
    if ((x == '^') || (x == '!')) {
      invert = true;
      x = need("unclosed inverted glob class");
    }
    if ((x == '-') || (x == ']')) {
      set.add(x);
      x = need("unclosed glob class");
    }
    while (x != ']') {
      uint8_t y = need("unclosed glob class");
      if (y == '-') {
        y = need("unended glob class range");
        if (y == ']') {
          set.add(x);
          set.add('-');
          x = y;
          continue;
        }
        if (x > y) fail_parse("backward glob class range", ptr - beg);
        set.addRange(x, y);
        x = need("unclosed glob class range");
      }
      else {
        set.add(x);
        x = y;
      }
    }
    if (invert)
      set.invert();
    return nfa.oneOf(set);
  }

  void addGlob(const uint8_t *p, size_t n, int32_t result, uint32_t f) {
    checkResult(result);
    nfa.setGoal(result);
    flags = f;
    NodeId startWild = (flags & RED_B200_F_LOOSE_START) ? nfa.wildcard() : 0;
    NodeId s = 0;
    const uint8_t *end = p + n;
    for (const uint8_t *ptr = p; ptr < end; ++ptr) {
      ByteSet one;
      switch (*ptr) {
      case '*': s = nfa.sequence(s, nfa.wildcard()); break;
      case '?': s = nfa.sequence(s, nfa.anyByte()); break;
      case '[': s = nfa.sequence(s, globClass(ptr, p, end)); break;
      default: one.add(*ptr); s = nfa.sequence(s, nfa.oneOf(one)); break;
      }
    }
    if (!s) fail_parse("uncategorized parsing error");
    finishPattern(s, startWild, result, false);
  }

  // byte-for-byte string (reference: Parser::addExact, lib/Parser.cpp:158):
  // a chain of nodes; with a tight end the end mark hangs directly off the
  // last node, with a loose end that node accepts before the wildcard is added
  void addExact(const uint8_t *p, size_t n, int32_t result, uint32_t f) {
    checkResult(result);
    nfa.setGoal(result);
    flags = f;
    NodeId startWild = (flags & RED_B200_F_LOOSE_START) ? nfa.wildcard() : 0;
    NodeId s = 0;
    NodeId cur = nfa.node(0);
    for (size_t i = 0; i < n; ++i) {
      NodeId nx = nfa.node(0);
      Edge e;
      e.to = nx;
      e.on.add(p[i]);
      nfa[cur].out.push_back(e);
      s = nfa.sequence(s, cur);
      cur = nx;
    }
    const bool loose = (flags & RED_B200_F_LOOSE_END) != 0;
    if (loose)
      nfa[cur].result = result;
    else {
      NodeId g = nfa.goalNode();
      Edge e;
      e.to = g;
      e.endMark = result;
      nfa[cur].out.push_back(e);
    }
    finishPattern(s, startWild, result, true);
  }
};

PatternParser::PatternParser(uint64_t maxStates, uint64_t maxParens)
    : impl_(new Impl(maxStates ? maxStates : ~0ull, maxParens ? maxParens : ~0ull)) {}

PatternParser::~PatternParser() { delete impl_; }

void PatternParser::add(int lang, const uint8_t *p, size_t n, int32_t result,
                        uint32_t flags) {
  switch (lang) {
  case RED_B200_LANG_RAW:
    impl_->addRegex(p, n, result, flags);
    break;
  case RED_B200_LANG_AUTO: {
    // grep-like heuristic (reference: Parser::addAuto, lib/Parser.cpp:92)
    flags |= RED_B200_F_LOOSE_START | RED_B200_F_LOOSE_END;
    if ((n >= 2) && (p[0] == '\\') && (p[1] == 'i')) {
      p += 2; n -= 2;
      flags |= RED_B200_F_IGNORE_CASE;
    }
    if ((n >= 1) && (p[0] == '^')) {
      p += 1; n -= 1;
      flags &= ~uint32_t(RED_B200_F_LOOSE_START);
    }
    else if ((n >= 2) && (p[0] == '.') && (p[1] == '*')) {
      p += 2; n -= 2;
      flags |= RED_B200_F_LOOSE_START;
    }
    if ((n >= 1) && (p[n - 1] == '$')) {
      n -= 1;
      flags &= ~uint32_t(RED_B200_F_LOOSE_END);
    }
    else if ((n >= 2) && (p[n - 2] == '.') && (p[n - 1] == '*')) {
      n -= 2;
      flags |= RED_B200_F_LOOSE_END;
    }
    impl_->addRegex(p, n, result, flags);
    break;
  }
  case RED_B200_LANG_GLOB:
    impl_->addGlob(p, n, result, flags);
    break;
  case RED_B200_LANG_EXACT:
    impl_->addExact(p, n, result, flags);
    break;
  default:
    fail_api("unrecognized regex language");
  }
}

Nfa &PatternParser::finish() {
  Nfa &nfa = impl_->nfa;
  if (nfa.size() == 0)
    nfa.setInitial(nfa.node(1));  // no patterns: the empty text matches, result 1
  nfa.pruneDeadEdges();
  return nfa;
}

} // namespace rx
} // namespace red_b200
