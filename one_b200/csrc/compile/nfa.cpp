/* nfa.cpp - see nfa.h. */

#include "nfa.h"

#include <algorithm>
#include <limits>

#include "../../../include/red_b200.h"

namespace red_b200 {
namespace rx {

void fail_parse(const std::string &msg) { throw CompileError{RED_B200_EPARSE, msg}; }
void fail_parse(const std::string &msg, size_t pos) {
  throw CompileError{RED_B200_EPARSE, msg + " near position " + std::to_string(pos)};
}
void fail_api(const std::string &msg) { throw CompileError{RED_B200_EAPI, msg}; }
void fail_limit(const std::string &msg) { throw CompileError{RED_B200_ELIMIT, msg}; }
void fail_internal(const std::string &msg) { throw CompileError{RED_B200_EINTERNAL, msg}; }

/* How the combinators are put together here.  What a combinator must DO to the
   automaton is fixed by the reference (nfa.h, parity note; lib/Nfa.cpp:160-262
   is where its effects were read off); how it is done is this file's:

     survey(x)      one traversal of a fragment: its nodes, and every edge that
                    enters a node which accepts at this moment;
     twin(edges, t) each surveyed edge gets a copy that goes to t instead;
     absorb(a, b)   a takes over b's outgoing edges (duplicates dropped).

   sequence = twin the entering edges to the right operand (+ absorb if the
   left one accepts at once) and stop the left side from accepting; star = twin
   them back to the start; alternate = absorb.  repeat() first takes all the
   copies it needs from the untouched operand and then chains them. */

// (id 0 is never a node; the reference charges its budget for the slot: the
// limits agree to the state)
NodeId Nfa::node(int32_t result) {
  if (nodes_.empty()) {
    nodes_.resize(1);
    used_ += 1;
  }
  if (nodes_.size() >= static_cast<size_t>(std::numeric_limits<NodeId>::max()))
    fail_limit("nfa state id overflow");
  if (used_ + 1 > budget_)
    fail_limit("state budget exceeded");
  used_ += 1;
  Node fresh;
  fresh.result = result;
  nodes_.push_back(std::move(fresh));
  return static_cast<NodeId>(nodes_.size() - 1);
}

void Nfa::addEdge(Node &n, const Edge &e) {
  if (std::find(n.out.begin(), n.out.end(), e) == n.out.end())
    n.out.push_back(e);
}

Nfa::Survey Nfa::survey(NodeId from) const {
  Survey sv;
  std::vector<char> seen(nodes_.size(), 0);
  std::vector<NodeId> stack(1, from);
  seen[from] = 1;
  while (!stack.empty()) {
    const NodeId id = stack.back();
    stack.pop_back();
    for (const Edge &e : nodes_[id].out)
      if (!seen[e.to]) {
        seen[e.to] = 1;
        stack.push_back(e.to);
      }
  }
  for (size_t id = 1; id < seen.size(); ++id) {
    if (!seen[id])
      continue;
    sv.nodes.push_back(static_cast<NodeId>(id));
    const std::vector<Edge> &out = nodes_[id].out;
    for (size_t k = 0; k < out.size(); ++k)
      if (nodes_[out[k].to].result > 0)
        sv.entering.push_back({static_cast<NodeId>(id), out[k]});
  }
  return sv;
}

void Nfa::twin(const Survey &sv, NodeId target) {
  for (const Survey::Entering &en : sv.entering) {
    Edge copy = en.edge;
    copy.to = target;
    addEdge(nodes_[en.from], copy);
  }
}

void Nfa::absorb(NodeId into, NodeId from) {
  const std::vector<Edge> taken = nodes_[from].out;  // (into may be from)
  for (const Edge &e : taken)
    addEdge(nodes_[into], e);
}

NodeId Nfa::fragment(const Edge &proto) {
  const NodeId start = node(0);
  Edge e = proto;
  e.to = goalNode();
  nodes_[start].out.push_back(e);
  return start;
}

NodeId Nfa::oneOf(const ByteSet &set) {
  Edge e;
  e.on = set;
  return fragment(e);
}

NodeId Nfa::anyByte() {
  Edge e;
  e.on.fill();
  return fragment(e);
}

NodeId Nfa::wildcard() { return star(anyByte()); }

NodeId Nfa::endMark(int32_t result) {
  Edge e;
  e.endMark = result;
  return fragment(e);
}

// x|y: x's start takes over y's first edges - whatever else points at that
// start (that is the reference's union, over-approximation included)
NodeId Nfa::alternate(NodeId x, NodeId y) {
  Node &left = nodes_[x];
  if ((left.result <= 0) && (nodes_[y].result > 0))
    left.result = nodes_[y].result;
  absorb(x, y);
  return x;
}

// xy: what entered an accepting node of x may enter y instead; x stops
// accepting - except that a start that accepted keeps doing so if y's does
NodeId Nfa::sequence(NodeId x, NodeId y) {
  if (!x || !y)
    return x ? x : y;
  const Survey sv = survey(x);
  const bool startAccepted = nodes_[x].result > 0;
  const int32_t kept = (startAccepted && (nodes_[y].result > 0)) ? nodes_[x].result : 0;
  twin(sv, y);
  if (startAccepted)
    absorb(x, y);
  for (NodeId id : sv.nodes)
    nodes_[id].result = 0;
  nodes_[x].result = kept;
  return x;
}

NodeId Nfa::optional(NodeId x) {
  nodes_[x].result = goal_;
  return x;
}

// x*: what entered an accepting node may go back to the start; the start accepts
NodeId Nfa::star(NodeId x) {
  twin(survey(x), x);
  return optional(x);
}

NodeId Nfa::chain(const std::vector<NodeId> &parts) {
  NodeId all = 0;
  for (NodeId part : parts)
    all = sequence(all, part);
  return all;
}

// x{min,max}; max < 0: no upper bound; min < 0: exactly max
NodeId Nfa::repeat(NodeId x, int min, int max) {
  if ((min == 0) && (max < 0))
    return star(x);
  if (min < 0)
    min = max;
  // every copy from the operand as it is now: it is the first link of the chain
  // and changes as soon as something is appended to it
  const int mandatoryCopies = (min > 0) ? min - 1 : 0;
  const int optionalCopies = (max < 0) ? 0 : (min > 0 ? max - min : max - 1);
  std::vector<NodeId> must, may;
  for (int i = 0; i < optionalCopies; ++i)
    may.push_back(clone(x));
  const NodeId loop = ((max < 0) && (min > 0)) ? clone(x) : 0;
  for (int i = 0; i < mandatoryCopies; ++i)
    must.push_back(clone(x));
  for (NodeId &id : may)
    id = optional(id);
  const NodeId rest = loop ? star(loop) : chain(may);
  if (min == 0)
    return sequence(optional(x), rest);
  return sequence(sequence(x, chain(must)), rest);
}

void ByteSet::foldAsciiCase() {
  // 'A'..'Z' are bits 1..26 of the second word, 'a'..'z' bits 33..58
  const uint64_t upper = w[1] & 0x0000000007fffffeull;
  const uint64_t lower = w[1] & 0x07fffffe00000000ull;
  w[1] |= (upper << 32) | (lower >> 32);
}

NodeId Nfa::foldCase(NodeId x) {
  for (NodeId id : survey(x).nodes)
    for (Edge &e : nodes_[id].out)
      e.on.foldAsciiCase();
  return x;
}

// a copy of everything reachable from x (the shape is what matters, not the numbering)
NodeId Nfa::clone(NodeId x) {
  if (!x)
    return x;
  const std::vector<NodeId> members = survey(x).nodes;
  std::vector<NodeId> copyOf(nodes_.size(), 0);
  for (NodeId id : members)
    copyOf[id] = node(nodes_[id].result);
  for (NodeId id : members) {
    std::vector<Edge> out = nodes_[id].out;
    for (Edge &e : out)
      e.to = copyOf[e.to];
    nodes_[copyOf[id]].out.swap(out);
  }
  return copyOf[x];
}

void Nfa::joinInitial(NodeId x) {
  init_ = init_ ? alternate(init_, x) : x;
}

// edges into nodes that neither accept nor lead anywhere go (one pass: a node
// that only now has no edges left stays, as in the reference), and so do
// repeated edges
void Nfa::pruneDeadEdges() {
  std::vector<char> dead(nodes_.size(), 0);
  for (size_t id = 1; id < nodes_.size(); ++id)
    dead[id] = (nodes_[id].result <= 0) && nodes_[id].out.empty();
  for (Node &n : nodes_) {
    std::vector<Edge> kept;
    kept.reserve(n.out.size());
    for (const Edge &e : n.out)
      if (!dead[e.to] && (std::find(kept.begin(), kept.end(), e) == kept.end()))
        kept.push_back(e);
    n.out.swap(kept);
  }
}

} // namespace rx
} // namespace red_b200
