/* parser.cpp - RED's pattern languages -> NFA.

   The dialect is the reference's (doc/Usage.md "Regular Expression Syntax";
   which inputs are accepted and what they mean was established against
   lib/Scanner.cpp and lib/Parser.cpp by the differential tests in
   tests/test_compile.py) - the implementation is not: a pattern is first
   turned into a flat token list, the tokens into a small syntax tree by an
   iterative operator-precedence pass, and the tree is emitted bottom-up
   through the combinators of nfa.h.  Globs are translated into the same
   tokens.  Error texts are this file's own.

   Dialect notes that the code below has to honour (all of them observable
   through the reference's accept/reject behaviour):
     * inside [...] a leading ']' or '-' is literal; "x-]" is the two bytes x
       and '-'; the ends of a range are taken as raw bytes even when written
       as \d or \w;
     * {n} needs n > 0; {,} is *, {0,} is *, {0,0} and {,0} are errors;
     * an atom may be followed by ONE repeat; \i may follow an atom or a
       repeat, but a repeat may not follow \i;
     * an alternative may be empty ("a|", "(|b)", "()").
 */

#include "parser.h"

#include <limits>
#include <vector>

#include "../../../include/red_b200.h"

namespace red_b200 {
namespace rx {

namespace {

// ---- bytes and escapes -----------------------------------------------------

// one decoded unit of pattern text
struct Unit {
  int value = -1;     // the byte; -1: end of pattern
  bool named = false; // written as backslash + a character without a fixed byte value (\d, \i, \q ...)
};

int hexDigit(uint8_t c) {
  if ((c >= '0') && (c <= '9')) return c - '0';
  c |= 0x20;
  if ((c >= 'a') && (c <= 'f')) return c - 'a' + 10;
  return -1;
}

class Reader {
public:
  Reader(const uint8_t *p, size_t n) : beg_(p), cur_(p), end_(p + n) {}
  size_t at() const { return static_cast<size_t>(cur_ - beg_); }
  bool done() const { return cur_ >= end_; }
  uint8_t peek() const { return *cur_; }
  uint8_t take() { return *cur_++; }

  Unit unit() {
    Unit u;
    if (done())
      return u;
    uint8_t c = take();
    if (c != '\\') {
      u.value = c;
      return u;
    }
    if (done())
      fail_parse("a backslash ends the pattern", at());
    c = take();
    static const char plain[] = "0\0a\ab\bn\nr\rt\tv\v\\\\";  // pairs: letter, byte
    for (size_t i = 0; i + 1 < sizeof(plain); i += 2)
      if (c == static_cast<uint8_t>(plain[i])) {
        u.value = static_cast<uint8_t>(plain[i + 1]);
        return u;
      }
    if (c == 'x') {
      const int hi = (end_ - cur_ >= 2) ? hexDigit(cur_[0]) : -1;
      const int lo = (end_ - cur_ >= 2) ? hexDigit(cur_[1]) : -1;
      if ((hi < 0) || (lo < 0))
        fail_parse("\\x wants two hexadecimal digits", at());
      cur_ += 2;
      u.value = hi * 16 + lo;
      return u;
    }
    u.value = c;
    u.named = true;
    return u;
  }

private:
  const uint8_t *beg_, *cur_, *end_;
};

// what a unit contributes to a set: \d \D \s \S \w \W are sets, the rest one byte
void unitInto(ByteSet &set, const Unit &u) {
  if (!u.named) {
    set.add(u.value);
    return;
  }
  ByteSet base;
  const int lower = u.value | 0x20;
  if (lower == 'd')
    base.addRange('0', '9');
  else if (lower == 's') {
    base.add(' ');
    base.addRange('\t', '\r');
  }
  else if (lower == 'w') {
    base.addRange('0', '9');
    base.addRange('A', 'Z');
    base.addRange('a', 'z');
    base.add('_');
  }
  else {
    set.add(u.value);  // \q is q
    return;
  }
  if (u.value != lower)
    base.invert();  // the capital letter is the complement
  set.merge(base);
}

// ---- tokens ------------------------------------------------------------------

enum class Tk { End, Fold, Set, Times, Or, Open, Close, Bad };

struct Token {
  Tk kind = Tk::End;
  size_t at = 0;
  ByteSet set;        // Set
  int lo = 0, hi = 0; // Times; hi < 0: no upper bound
  const char *why = nullptr;  // Bad: reported when the parser gets there
};

Token bad(size_t at, const char *why) {
  Token t;
  t.kind = Tk::Bad;
  t.at = at;
  t.why = why;
  return t;
}

/* [...]: the units up to the closing bracket are decoded first, then read with
   two units of lookahead ("x-y" is a range, "x-]" is not). */
Token bracketSet(Reader &in, bool regexEscapes) {
  Token t;
  t.kind = Tk::Set;
  t.at = in.at();
  std::vector<Unit> units;
  auto get = [&](size_t i) -> const Unit & {
    while (units.size() <= i) {
      Unit u;
      if (regexEscapes)
        u = in.unit();
      else if (!in.done())
        u.value = in.take();
      units.push_back(u);
    }
    return units[i];
  };
  auto is = [&](size_t i, int c) { return !get(i).named && (get(i).value == c); };
  size_t i = 0;
  bool negate = false;
  if (is(0, '^') || (!regexEscapes && is(0, '!'))) {
    negate = true;
    i = 1;
  }
  if (is(i, ']') || is(i, '-'))
    t.set.add(get(i++).value);  // literal in first place
  for (;;) {
    if (get(i).value < 0)
      return bad(in.at(), "a character class is not closed");
    if (is(i, ']'))
      break;
    if (is(i + 1, '-')) {
      if (get(i + 2).value < 0)
        return bad(in.at(), "a character class ends inside a range");
      if (is(i + 2, ']')) {
        t.set.add(get(i).value);
        t.set.add('-');
        break;
      }
      const int from = get(i).value, to = get(i + 2).value;  // raw bytes, also for \d and the like
      if (from > to)
        return bad(in.at(), "a character class range runs downwards");
      t.set.addRange(from, to);
      i += 3;
      continue;
    }
    if (regexEscapes)
      unitInto(t.set, get(i));
    else
      t.set.add(get(i).value);
    ++i;
  }
  if (negate)
    t.set.invert();
  return t;
}

/* {n} {n,m} {n,} {,m} {,} */
Token braceTimes(Reader &in) {
  Token t;
  t.kind = Tk::Times;
  t.at = in.at();
  long value[2] = {0, 0};
  bool digits[2] = {false, false};
  int field = 0;
  for (;;) {
    if (in.done())
      return bad(in.at(), "a repeat count is not closed");
    const uint8_t c = in.take();
    if (c == '}')
      break;
    if (c == ',') {
      if (field == 1)
        return bad(in.at(), "a repeat count has more than one comma");
      field = 1;
    }
    else if ((c >= '0') && (c <= '9')) {
      value[field] = value[field] * 10 + (c - '0');
      digits[field] = true;
    }
    else
      return bad(in.at(), "a repeat count may hold digits and one comma only");
  }
  if (field == 0) {
    if (!digits[0] || (value[0] == 0))
      return bad(in.at(), "a repeat count must be positive");
    t.lo = t.hi = static_cast<int>(value[0]);
    return t;
  }
  t.lo = static_cast<int>(value[0]);
  t.hi = digits[1] ? static_cast<int>(value[1]) : -1;
  if (digits[1] && (value[1] == 0) && (value[0] == 0))
    return bad(in.at(), "a repeat count of zero");
  if (digits[1] && (value[1] < value[0]))
    return bad(in.at(), "a repeat count's upper bound is below its lower");
  return t;
}

std::vector<Token> regexTokens(const uint8_t *p, size_t n) {
  std::vector<Token> out;
  Reader in(p, n);
  auto simple = [&](Tk kind) {
    Token t;
    t.kind = kind;
    t.at = in.at();
    out.push_back(t);
  };
  auto times = [&](int lo, int hi) {
    Token t;
    t.kind = Tk::Times;
    t.at = in.at();
    t.lo = lo;
    t.hi = hi;
    out.push_back(t);
  };
  while (!in.done() && (out.empty() || (out.back().kind != Tk::Bad))) {
    switch (in.peek()) {
    case '(': in.take(); simple(Tk::Open); continue;
    case ')': in.take(); simple(Tk::Close); continue;
    case '|': in.take(); simple(Tk::Or); continue;
    case '*': in.take(); times(0, -1); continue;
    case '+': in.take(); times(1, -1); continue;
    case '?': in.take(); times(0, 1); continue;
    case '[': in.take(); out.push_back(bracketSet(in, true)); continue;
    case '{': in.take(); out.push_back(braceTimes(in)); continue;
    default: break;
    }
    Token t;
    t.at = in.at();
    if (in.peek() == '.') {
      in.take();
      t.kind = Tk::Set;
      t.set.fill();
      out.push_back(t);
      continue;
    }
    const Unit u = in.unit();
    if (u.named && (u.value >= '1') && (u.value <= '9')) {
      out.push_back(bad(t.at, "backreferences cannot be matched by a DFA"));
      break;
    }
    if (u.named && (u.value == 'i'))
      t.kind = Tk::Fold;
    else {
      t.kind = Tk::Set;
      unitInto(t.set, u);
    }
    out.push_back(t);
  }
  Token end;
  end.at = n;
  if (out.empty() || (out.back().kind != Tk::Bad))
    out.push_back(end);
  return out;
}

/* A glob as the same tokens: '*' is a Set of everything under a star that is
   marked as a wildcard (lo = hi = -2, emitted as the NFA's own ".*" fragment),
   '?' one byte of any value, [...] a class without escapes ('!' negates as
   well), anything else itself. */
std::vector<Token> globTokens(const uint8_t *p, size_t n) {
  std::vector<Token> out;
  Reader in(p, n);
  while (!in.done() && (out.empty() || (out.back().kind != Tk::Bad))) {
    Token t;
    t.at = in.at();
    const uint8_t c = in.take();
    if (c == '[') {
      out.push_back(bracketSet(in, false));
      continue;
    }
    t.kind = Tk::Set;
    if (c == '*')
      t.lo = t.hi = -2;
    else if (c == '?')
      t.set.fill();
    else
      t.set.add(c);
    out.push_back(t);
  }
  Token end;
  end.at = n;
  if (out.empty() || (out.back().kind != Tk::Bad))
    out.push_back(end);
  return out;
}

// ---- syntax tree ------------------------------------------------------------

struct Tree {
  enum Kind { Nothing, Bytes, Wild, Seq, Alt, Times };
  struct Node {
    Kind kind = Nothing;
    ByteSet set;             // Bytes
    int lo = 0, hi = 0;      // Times
    int left = -1, right = -1;
  };
  std::vector<Node> nodes;
  int add(Kind k, int l = -1, int r = -1) {
    Node n;
    n.kind = k;
    n.left = l;
    n.right = r;
    nodes.push_back(n);
    return static_cast<int>(nodes.size()) - 1;
  }
};

/* tokens -> tree, without recursion: a stack of open groups, each holding the
   alternatives finished so far and the sequence being collected. */
class Builder {
public:
  Builder(const std::vector<Token> &toks, uint64_t nesting) : toks_(toks), nestingLeft_(nesting) {}

  int run(Tree &tree, bool &fold) {
    std::vector<Frame> open(1);
    size_t i = 0;
    for (;;) {
      const Token &t = toks_[i];
      Frame &f = open.back();
      switch (t.kind) {
      case Tk::Bad:
        fail_parse(t.why, t.at);
      case Tk::Fold:
        fold = true;
        if (f.lastAtom >= 0)
          f.afterFold = true;  // a repeat may not come next
        ++i;
        break;
      case Tk::Set:
        push(tree, f, (t.lo == -2) ? tree.add(Tree::Wild) : bytes(tree, t.set));
        ++i;
        break;
      case Tk::Times: {
        if ((f.lastAtom < 0) || f.afterFold || f.repeated)
          fail_parse("a repeat has nothing to repeat", t.at);
        const int inner = f.lastAtom;
        const int rep = tree.add(Tree::Times, inner);
        tree.nodes[rep].lo = t.lo;
        tree.nodes[rep].hi = t.hi;
        replaceLast(tree, f, rep);
        f.repeated = true;
        ++i;
        break;
      }
      case Tk::Or:
        f.alts.push_back(f.seq);
        f.seq = -1;
        f.lastAtom = -1;
        f.afterFold = f.repeated = false;
        ++i;
        break;
      case Tk::Open:
        if (nestingLeft_ == 0)
          fail_limit("parentheses are nested deeper than the budget allows");
        --nestingLeft_;
        open.emplace_back();
        ++i;
        break;
      case Tk::Close: {
        if (open.size() == 1)
          fail_parse("a closing parenthesis without an opening one", t.at);
        const int group = close(tree, f);
        open.pop_back();
        ++nestingLeft_;
        push(tree, open.back(), group);
        ++i;
        break;
      }
      case Tk::End:
        if (open.size() != 1)
          fail_parse("a parenthesis is not closed", t.at);
        return close(tree, f);
      }
    }
  }

private:
  struct Frame {
    std::vector<int> alts;  // finished alternatives (-1: empty)
    int seq = -1;           // the sequence being collected (-1: nothing yet)
    int lastAtom = -1;      // its last element
    int lastParent = -1;    // the Seq node whose right child lastAtom is (-1: it is seq itself)
    bool afterFold = false, repeated = false;
  };

  static int bytes(Tree &tree, const ByteSet &set) {
    const int n = tree.add(Tree::Bytes);
    tree.nodes[n].set = set;
    return n;
  }
  static void push(Tree &tree, Frame &f, int atom) {
    if (f.seq < 0) {
      f.seq = atom;
      f.lastParent = -1;
    }
    else {
      f.seq = tree.add(Tree::Seq, f.seq, atom);  // left-leaning: ((a b) c) d
      f.lastParent = f.seq;
    }
    f.lastAtom = atom;
    f.afterFold = f.repeated = false;
  }
  static void replaceLast(Tree &tree, Frame &f, int with) {
    if (f.lastParent < 0)
      f.seq = with;
    else
      tree.nodes[f.lastParent].right = with;
    f.lastAtom = with;
  }
  // the group's alternatives, left to right: ((a|b)|c)
  static int close(Tree &tree, Frame &f) {
    f.alts.push_back(f.seq);
    int all = -2;
    for (int a : f.alts) {
      const int one = (a < 0) ? tree.add(Tree::Nothing) : a;
      all = (all == -2) ? one : tree.add(Tree::Alt, all, one);
    }
    return all;
  }

  const std::vector<Token> &toks_;
  uint64_t nestingLeft_;
};

// tree -> NFA fragment, children first (left before right, as the patterns read)
NodeId emit(Nfa &nfa, const Tree &tree, int root) {
  // post-order without recursion
  std::vector<NodeId> made(tree.nodes.size(), 0);
  std::vector<std::pair<int, bool>> todo{{root, false}};
  while (!todo.empty()) {
    const int n = todo.back().first;
    const bool ready = todo.back().second;
    todo.pop_back();
    const Tree::Node &node = tree.nodes[n];
    if (!ready) {
      todo.push_back({n, true});
      if (node.right >= 0)
        todo.push_back({node.right, false});
      if (node.left >= 0)
        todo.push_back({node.left, false});
      continue;
    }
    switch (node.kind) {
    case Tree::Nothing: made[n] = nfa.goalNode(); break;  // the empty text
    case Tree::Bytes: made[n] = nfa.oneOf(node.set); break;
    case Tree::Wild: made[n] = nfa.wildcard(); break;
    case Tree::Seq: made[n] = nfa.sequence(made[node.left], made[node.right]); break;
    case Tree::Alt: made[n] = nfa.alternate(made[node.left], made[node.right]); break;
    case Tree::Times: made[n] = nfa.repeat(made[node.left], node.lo, node.hi); break;
    }
  }
  return made[root];
}

// prefixes and suffixes addAuto strips, and what they do to the flags
struct Affix {
  const char *text;
  uint32_t set, clear;
};
const Affix kAutoHead[] = {{"^", 0, RED_B200_F_LOOSE_START}, {".*", RED_B200_F_LOOSE_START, 0}};
const Affix kAutoTail[] = {{"$", 0, RED_B200_F_LOOSE_END}, {".*", RED_B200_F_LOOSE_END, 0}};

bool startsWith(const uint8_t *p, size_t n, const char *s) {
  size_t k = 0;
  while (s[k] && (k < n) && (p[k] == static_cast<uint8_t>(s[k])))
    ++k;
  return s[k] == 0;
}
bool endsWith(const uint8_t *p, size_t n, const char *s) {
  size_t m = 0;
  while (s[m])
    ++m;
  return (n >= m) && startsWith(p + (n - m), m, s);
}

} // anonymous

struct PatternParser::Impl {
  Nfa nfa;
  uint64_t nesting;

  Impl(uint64_t states, uint64_t parens) : nfa(states), nesting(parens) {}

  /* A finished fragment becomes a rule: case folding, a ".*" in front and/or
     behind for the loose ends, the end mark that carries the result id to the
     subset construction, and the join with the rules added before.
     markInside: the fragment already ends in its end mark (exact strings). */
  void install(NodeId body, int32_t result, uint32_t flags, bool markInside) {
    NodeId front = (flags & RED_B200_F_LOOSE_START) ? nfa.wildcard() : 0;
    if (flags & RED_B200_F_IGNORE_CASE)
      body = nfa.foldCase(body);
    if (front)
      body = nfa.sequence(front, body);
    if (flags & RED_B200_F_LOOSE_END)
      body = nfa.sequence(body, nfa.wildcard());
    if (!markInside || (flags & RED_B200_F_LOOSE_END))
      body = nfa.sequence(body, nfa.endMark(result));
    nfa.joinInitial(body);
  }

  void begin(int32_t result) {
    if (result <= 0)
      fail_api("a result id must be positive");
    if (result > std::numeric_limits<int32_t>::max() - 256)
      fail_api("a result id this large is reserved");
    nfa.setGoal(result);
  }

  void fromTokens(const std::vector<Token> &toks, int32_t result, uint32_t flags) {
    begin(result);
    // (the reference makes the leading wildcard before it parses: keep the
    //  node order, it shows in nothing but makes traces comparable)
    Tree tree;
    bool fold = false;
    Builder builder(toks, nesting);
    const int root = builder.run(tree, fold);
    if (fold)
      flags |= RED_B200_F_IGNORE_CASE;
    install(emit(nfa, tree, root), result, flags, false);
  }

  /* Byte for byte: a chain of nodes, one per byte.  With a tight end the end
     mark hangs directly off the last node; with a loose end that node accepts
     and the wildcard and the mark follow (install). */
  void exact(const uint8_t *p, size_t n, int32_t result, uint32_t flags) {
    begin(result);
    const bool loose = (flags & RED_B200_F_LOOSE_END) != 0;
    std::vector<NodeId> chain(n + 1);
    for (NodeId &id : chain)
      id = nfa.node(0);
    NodeId body = 0;
    for (size_t i = 0; i < n; ++i) {
      Edge e;
      e.to = chain[i + 1];
      e.on.add(p[i]);
      nfa[chain[i]].out.push_back(e);
      body = nfa.sequence(body, chain[i]);
    }
    if (loose)
      nfa[chain[n]].result = result;
    else {
      Edge e;
      e.to = nfa.goalNode();
      e.endMark = result;
      nfa[chain[n]].out.push_back(e);
    }
    install(body, result, flags, true);
  }
};

PatternParser::PatternParser(uint64_t maxStates, uint64_t maxParens)
    : impl_(new Impl(maxStates ? maxStates : ~0ull, maxParens ? maxParens : ~0ull)) {}

PatternParser::~PatternParser() { delete impl_; }

void PatternParser::add(int lang, const uint8_t *p, size_t n, int32_t result,
                        uint32_t flags) {
  switch (lang) {
  case RED_B200_LANG_AUTO:
    // what grep users write: loose at both ends unless anchored with ^ / $,
    // an explicit .* at either end is the same as no anchor, a leading \i
    // folds case
    flags |= RED_B200_F_LOOSE_START | RED_B200_F_LOOSE_END;
    if (startsWith(p, n, "\\i")) {
      p += 2;
      n -= 2;
      flags |= RED_B200_F_IGNORE_CASE;
    }
    for (const Affix &a : kAutoHead)
      if (startsWith(p, n, a.text)) {
        const size_t m = (a.text[1] ? 2 : 1);
        p += m;
        n -= m;
        flags = (flags | a.set) & ~a.clear;
        break;
      }
    for (const Affix &a : kAutoTail)
      if (endsWith(p, n, a.text)) {
        n -= (a.text[1] ? 2 : 1);
        flags = (flags | a.set) & ~a.clear;
        break;
      }
    impl_->fromTokens(regexTokens(p, n), result, flags);
    break;
  case RED_B200_LANG_RAW:
    impl_->fromTokens(regexTokens(p, n), result, flags);
    break;
  case RED_B200_LANG_GLOB:
    if (n == 0)
      fail_parse("an empty glob");
    impl_->fromTokens(globTokens(p, n), result, flags);
    break;
  case RED_B200_LANG_EXACT:
    impl_->exact(p, n, result, flags);
    break;
  default:
    fail_api("unknown pattern language");
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
