/* nfa.h - epsilon-free NFA with the construction semantics of RED.

   Host-side C++.  This is the first stage of the compile pipeline
   (patterns -> NFA -> DFA -> minimal DFA -> serialized program).

   Parity note.  A pattern compiled here must match exactly what the same
   pattern compiled by the reference matches - including where the reference's
   construction is not textbook.  Its combinators work on epsilon-free
   automata by copying transitions (reference: lib/Nfa.cpp:160-262), and some
   of them over-approximate: union copies the right operand's initial
   transitions into the left operand's initial state even when that state is a
   loop target, so `a*|b` accepts "ab" while `b|a*` does not.  Users' rule sets
   were written against that behaviour, so each combinator below is specified
   by its effect on the automaton exactly as the reference defines it, and the
   differential tests (tests/test_compile.py) pin the resulting DFAs to the
   reference's, state for state up to renumbering.

   Representation is our own: a transition is (target, 256-bit byte set,
   optional end mark).  End marks (reference: include/Powerset.h:22-29) are
   out-of-alphabet symbols that carry a pattern's result id to the subset
   construction; here they are a separate field rather than bits >= 256.
 */

#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace red_b200 {
namespace rx {

struct CompileError {
  int code;  // RED_B200_E*
  std::string msg;
};

[[noreturn]] void fail_parse(const std::string &msg);
[[noreturn]] void fail_parse(const std::string &msg, size_t pos);
[[noreturn]] void fail_api(const std::string &msg);
[[noreturn]] void fail_limit(const std::string &msg);
[[noreturn]] void fail_internal(const std::string &msg);

struct ByteSet {
  uint64_t w[4] = {0, 0, 0, 0};
  void add(unsigned c) { w[c >> 6] |= 1ull << (c & 63); }
  void addRange(unsigned lo, unsigned hi) {
    for (unsigned c = lo; c <= hi; ++c)
      add(c);
  }
  bool has(unsigned c) const { return (w[c >> 6] >> (c & 63)) & 1; }
  void invert() {
    for (uint64_t &x : w)
      x = ~x;
  }
  void fill() {
    for (uint64_t &x : w)
      x = ~0ull;
  }
  void merge(const ByteSet &o) {
    for (int i = 0; i < 4; ++i)
      w[i] |= o.w[i];
  }
  bool empty() const { return !(w[0] | w[1] | w[2] | w[3]); }
  void foldAsciiCase();  // every ASCII letter in the set brings its other case along
  bool operator==(const ByteSet &o) const {
    return (w[0] == o.w[0]) && (w[1] == o.w[1]) && (w[2] == o.w[2]) && (w[3] == o.w[3]);
  }
};

typedef int32_t NodeId;  // 0 = none

struct Edge {
  NodeId to = 0;
  ByteSet on;
  int32_t endMark = 0;  // > 0: the end mark of result `endMark` (no bytes)
  bool operator==(const Edge &o) const {
    return (to == o.to) && (endMark == o.endMark) && (on == o.on);
  }
};

struct Node {
  int32_t result = 0;  // > 0 while the node counts as accepting during construction
  std::vector<Edge> out;
};

class Nfa {
public:
  explicit Nfa(uint64_t stateBudget) : budget_(stateBudget) {}

  size_t size() const { return nodes_.size(); }
  const Node &operator[](NodeId id) const { return nodes_[id]; }
  Node &operator[](NodeId id) { return nodes_[id]; }

  NodeId initial() const { return init_; }
  void setInitial(NodeId id) { init_ = id; }
  void setGoal(int32_t g) { goal_ = g; }
  uint64_t statesUsed() const { return used_; }

  NodeId node(int32_t result);
  NodeId goalNode() { return node(goal_); }

  // building blocks (each returns the initial node of the fragment)
  NodeId oneOf(const ByteSet &set);   // a single byte from the set
  NodeId anyByte();                   // .
  NodeId wildcard();                  // .*
  NodeId endMark(int32_t result);

  // combinators, with the reference's semantics (see header comment; the
  // reference's name for the same effect in brackets)
  NodeId alternate(NodeId x, NodeId y);        // x|y      [stateUnion]
  NodeId sequence(NodeId x, NodeId y);         // xy       [stateConcat]
  NodeId optional(NodeId x);                   // x?       [stateOptional]
  NodeId star(NodeId x);                       // x*       [stateKleenStar]
  NodeId repeat(NodeId x, int min, int max);   // x{m,n}   [stateClosure]
  NodeId foldCase(NodeId x);                   // \i       [stateIgnoreCase]
  NodeId clone(NodeId x);                      //          [deepCopyState]

  void joinInitial(NodeId x);                  // one more rule of the union [selfUnion]
  void pruneDeadEdges();                       // [dropUselessTransitions]

private:
  // a fragment at one moment: its nodes (ascending) and the edges that enter
  // a node which accepts right now
  struct Survey {
    struct Entering { NodeId from; Edge edge; };
    std::vector<NodeId> nodes;
    std::vector<Entering> entering;
  };
  Survey survey(NodeId from) const;
  void twin(const Survey &sv, NodeId target);
  void absorb(NodeId into, NodeId from);
  NodeId fragment(const Edge &proto);            // start --proto--> goal
  NodeId chain(const std::vector<NodeId> &parts);
  void addEdge(Node &n, const Edge &e);

  std::vector<Node> nodes_;
  NodeId init_ = 0;
  int32_t goal_ = 0;
  uint64_t budget_;
  uint64_t used_ = 0;
};

} // namespace rx
} // namespace red_b200
