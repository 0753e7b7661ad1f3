/* nfa.cpp - see nfa.h.  Each combinator states the reference function whose
   effect on the automaton it reproduces (reference: lib/Nfa.cpp). */

#include "nfa.h"

#include <deque>
#include <limits>
#include <unordered_map>

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

// reference: NfaObj::newState (lib/Nfa.cpp:95) - id 0 is never a node
NodeId Nfa::node(int32_t result) {
  if (nodes_.empty())
    nodes_.resize(1);
  if (nodes_.size() >= static_cast<size_t>(std::numeric_limits<NodeId>::max()))
    fail_limit("nfa state id overflow");
  uint64_t cost = (nodes_.size() == 1) ? 2 : 1;
  if (used_ + cost > budget_)
    fail_limit("state budget exceeded");
  used_ += cost;
  nodes_.emplace_back();
  nodes_.back().result = result;
  return static_cast<NodeId>(nodes_.size() - 1);
}

void Nfa::addEdge(Node &n, const Edge &e) {
  for (const Edge &have : n.out)
    if (have == e)
      return;
  n.out.push_back(e);
}

// nodes reachable from `from` (itself included), ascending ids
std::vector<NodeId> Nfa::reachable(NodeId from) const {
  std::vector<char> seen(nodes_.size(), 0);
  std::deque<NodeId> todo;
  seen[from] = 1;
  todo.push_back(from);
  while (!todo.empty()) {
    NodeId id = todo.front();
    todo.pop_front();
    for (const Edge &e : nodes_[id].out)
      if (!seen[e.to]) {
        seen[e.to] = 1;
        todo.push_back(e.to);
      }
  }
  std::vector<NodeId> out;
  for (size_t i = 1; i < seen.size(); ++i)
    if (seen[i])
      out.push_back(static_cast<NodeId>(i));
  return out;
}

NodeId Nfa::oneOf(const ByteSet &set) {
  NodeId a = node(0);
  NodeId b = goalNode();
  Edge e;
  e.to = b;
  e.on = set;
  nodes_[a].out.push_back(e);
  return a;
}

NodeId Nfa::anyByte() {
  ByteSet all;
  all.fill();
  return oneOf(all);
}

NodeId Nfa::wildcard() { return star(anyByte()); }

NodeId Nfa::endMark(int32_t result) {
  NodeId a = node(0);
  NodeId b = goalNode();
  Edge e;
  e.to = b;
  e.endMark = result;
  nodes_[a].out.push_back(e);
  return a;
}

// reference: stateUnion (lib/Nfa.cpp:160) - y's initial edges are copied into
// x's initial node, whatever else points at that node
NodeId Nfa::alternate(NodeId x, NodeId y) {
  if ((nodes_[x].result <= 0) && (nodes_[y].result > 0))
    nodes_[x].result = nodes_[y].result;
  const std::vector<Edge> copy = nodes_[y].out;  // x may be y
  for (const Edge &e : copy)
    addEdge(nodes_[x], e);
  return x;
}

// reference: stateConcat (lib/Nfa.cpp:171) - every edge that enters an
// accepting node of x gets a twin that enters y; x's accepting nodes stop
// accepting (x itself keeps accepting only if y accepts too)
NodeId Nfa::sequence(NodeId x, NodeId y) {
  if (!x)
    return y;
  if (!y)
    return x;
  struct Pending { NodeId from; Edge e; };
  std::vector<NodeId> accepting;
  std::vector<Pending> entering;
  for (NodeId id : reachable(x)) {
    if (nodes_[id].result > 0)
      accepting.push_back(id);
    for (const Edge &e : nodes_[id].out)
      if (nodes_[e.to].result > 0)
        entering.push_back({id, e});
  }
  for (Pending &p : entering) {
    Edge twin = p.e;
    twin.to = y;
    addEdge(nodes_[p.from], twin);
  }
  const bool xAccepts = nodes_[x].result > 0;
  const bool yAccepts = nodes_[y].result > 0;
  if (xAccepts) {
    const std::vector<Edge> copy = nodes_[y].out;
    for (const Edge &e : copy)
      addEdge(nodes_[x], e);
  }
  int32_t keep = (xAccepts && yAccepts) ? nodes_[x].result : 0;
  for (NodeId id : accepting)
    nodes_[id].result = 0;
  if (keep > 0)
    nodes_[x].result = keep;
  return x;
}

// reference: stateOptional (lib/Nfa.cpp:208)
NodeId Nfa::optional(NodeId x) {
  nodes_[x].result = goal_;
  return x;
}

// reference: stateKleenStar (lib/Nfa.cpp:214) - edges entering an accepting
// node get a twin back to the start
NodeId Nfa::star(NodeId x) {
  struct Pending { NodeId from; Edge e; };
  std::vector<Pending> entering;
  for (NodeId id : reachable(x))
    for (const Edge &e : nodes_[id].out)
      if (nodes_[e.to].result > 0)
        entering.push_back({id, e});
  for (Pending &p : entering) {
    Edge twin = p.e;
    twin.to = x;
    addEdge(nodes_[p.from], twin);
  }
  return optional(x);
}

// reference: stateClosure (lib/Nfa.cpp:224) - x{min,max}; max < 0 = unbounded.
// All copies are taken from x before x itself is modified.
NodeId Nfa::repeat(NodeId x, int min, int max) {
  if (min == 0) {
    if (max < 0)
      return star(x);
    NodeId opt = 0;
    for (int i = 1; i < max; ++i)
      opt = sequence(opt, optional(clone(x)));
    return sequence(optional(x), opt);
  }
  if (min < 0)
    min = max;
  NodeId tail = 0;
  if (max < 0)
    tail = star(clone(x));
  else
    for (int i = 0; i < max - min; ++i)
      tail = sequence(tail, optional(clone(x)));
  NodeId head = 0;
  for (int i = 1; i < min; ++i)
    head = sequence(head, clone(x));
  head = sequence(x, head);
  return sequence(head, tail);
}

// reference: stateIgnoreCase (lib/Nfa.cpp:252)
NodeId Nfa::foldCase(NodeId x) {
  for (NodeId id : reachable(x))
    for (Edge &e : nodes_[id].out) {
      for (unsigned c = 'A'; c <= 'Z'; ++c)
        if (e.on.has(c))
          e.on.add(c + 32);
      for (unsigned c = 'a'; c <= 'z'; ++c)
        if (e.on.has(c))
          e.on.add(c - 32);
    }
  return x;
}

// reference: deepCopyState (lib/Nfa.cpp:111) - copy of everything reachable
NodeId Nfa::clone(NodeId x) {
  if (!x)
    return x;
  std::unordered_map<NodeId, NodeId> twin;
  // iterative preorder copy; the shape is what matters, not the numbering
  std::vector<NodeId> order = reachable(x);
  for (NodeId id : order)
    twin[id] = node(nodes_[id].result);
  for (NodeId id : order) {
    std::vector<Edge> out = nodes_[id].out;
    for (Edge &e : out)
      e.to = twin[e.to];
    nodes_[twin[id]].out = std::move(out);
  }
  return twin[x];
}

// reference: selfUnion (lib/Nfa.cpp:305)
void Nfa::joinInitial(NodeId x) {
  init_ = init_ ? alternate(init_, x) : x;
}

// reference: dropUselessTransitions (lib/Nfa.cpp:313) - one pass, as there
void Nfa::pruneDeadEdges() {
  std::vector<char> useless(nodes_.size(), 0);
  for (size_t id = 1; id < nodes_.size(); ++id)
    if ((nodes_[id].result <= 0) && nodes_[id].out.empty())
      useless[id] = 1;
  for (Node &n : nodes_) {
    std::vector<Edge> kept;
    for (const Edge &e : n.out) {
      if (useless[e.to])
        continue;
      bool dup = false;
      for (const Edge &k : kept)
        if (k == e) {
          dup = true;
          break;
        }
      if (!dup)
        kept.push_back(e);
    }
    n.out.swap(kept);
  }
}

} // namespace rx
} // namespace red_b200
