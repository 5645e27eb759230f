// ORACLE — TEST INFRASTRUCTURE ONLY (see score.hpp header).
//
// CPU restatement of the reference's score director, moves, candidate order and the
// local-search candidate loop around it.
//   ScoreDirector            solverforge-scoring/src/director/score_director/incremental.rs:141-224
//   ChangeMove               solverforge-solver/src/heuristic/move/change.rs:125-175
//   SwapMove                 heuristic/move/swap.rs:140-215
//   CompoundScalarMove       heuristic/move/compound_scalar.rs:245-319,397-406
//   ListChangeMove           heuristic/move/list_kernel/change.rs:23-153
//   ListSwapMove             heuristic/move/list_kernel/swap.rs:31-110
//   ListReverseMove          heuristic/move/list_kernel/reverse.rs:21-58 (2-opt segment reversal)
//   SublistChangeMove        heuristic/move/list_kernel/sublist_change.rs:17-125 + segment_layout.rs:49-75
//   SublistSwapMove          heuristic/move/list_kernel/sublist_swap.rs:17-170 + segment_layout.rs:118-165
//   KOptMove (one list)      heuristic/move/list_kernel/k_opt.rs:11-110 + k_opt_reconnection.rs (patterns); CPU only —
//                            groundwork for a device path, not yet scored on the GPU
//   evaluate_candidate       phase/localsearch/evaluation.rs:20-115
//   MoveStreamContext        heuristic/selector/move_selector/iter.rs:14-207
//   ChangeMove order         heuristic/selector/move_selector/change.rs:66-104,246-307
//   nearby list change order heuristic/selector/list_kernel/nearby_change.rs:102-232,
//                            heuristic/selector/nearby_list_support.rs:3-34
//   foragers                 phase/localsearch/forager.rs:70-425, forager/improving.rs
//   acceptors                phase/localsearch/acceptor/{hill_climbing,late_acceptance,simulated_annealing}.rs
//   candidate loop           phase/localsearch/phase/candidates.rs:47-285
#pragma once
#include <cmath>
#include <functional>
#include <limits>

#include "constraints.hpp"

namespace sfo {

using OptVal = std::optional<size_t>;

// Accessors a model provides (the reference generates these with macros).
template <class S>
struct Access {
  OptVal (*get)(const S&, size_t desc, size_t entity) = nullptr;
  void (*set)(S&, size_t desc, size_t entity, OptVal v) = nullptr;
  std::vector<size_t>& (*list)(S&, size_t desc, size_t entity) = nullptr;
  size_t (*entity_count)(const S&, size_t desc) = nullptr;
  void (*update_entity_shadows)(S&, size_t desc, size_t entity) = nullptr;  // domain/traits.rs:61-68
};

template <class S, class Sc>
struct ScoreDirector {
  S working;
  ConstraintSet<S, Sc> constraints;
  Access<S> access;
  Sc cached = Sc::zero();
  bool initialized = false;
  uint64_t score_calculations = 0;

  Sc calculate_score() {
    if (!initialized) {
      cached = constraints.initialize_all(working);
      initialized = true;
    }
    return cached;
  }
  Sc fresh_score() const { return constraints.evaluate_all(working); }
  void before_variable_changed(size_t desc, size_t entity) {
    if (!initialized) return;
    cached = cached + constraints.on_retract_all(working, entity, desc);
  }
  void after_variable_changed(size_t desc, size_t entity) {
    if (!initialized) return;
    if (access.update_entity_shadows) access.update_entity_shadows(working, desc, entity);
    cached = cached + constraints.on_insert_all(working, entity, desc);
  }
  struct ScoreState {
    Sc committed;
    bool initialized;
  };
  ScoreState snapshot_score_state() const { return {cached, initialized}; }
  void restore_score_state(ScoreState st) {
    if (st.initialized) {
      cached = st.committed;
      initialized = true;
    } else {
      constraints.reset_all();
      cached = Sc::zero();
      initialized = false;
    }
  }
  void reset() {
    constraints.reset_all();
    initialized = false;
    cached = Sc::zero();
  }
};

struct ScalarEdit {  // planning/scalar/candidate.rs:6-12
  size_t descriptor_index, entity_index;
  OptVal to_value;
};

struct Move {
  enum Kind { Change, Swap, Compound, ListChange, ListSwap, ListReverse, SublistChange, SublistSwap, KOpt } kind = Change;
  size_t desc = 0;
  size_t a = 0, b = 0, c = 0, d = 0;  // Change: a=entity; Swap: a,b; List*: a=src_e b=src_p c=dst_e d=dst_p
  size_t e = 0;                       // SublistChange: a=src_e [b, c)=source range d=dst_e e=dst_position
  size_t f = 0;                       // SublistSwap: a=first_e [b, c)  d=second_e [e, f)
  std::vector<size_t> cuts;           // KOpt: a=entity, k cut positions; the k + 1 segments are re-ordered by
  uint8_t seg_order[6] = {0, 0, 0, 0, 0, 0};  //   seg_order (position -> original segment) and reversed where
  uint8_t reverse_mask = 0, seg_len = 0;      //   bit i of reverse_mask is set (k_opt_reconnection.rs)
  OptVal to;
  std::vector<ScalarEdit> edits;
  bool requires_hard_improvement = false;
  bool requires_score_improvement = false;

  static Move change(size_t desc, size_t e, OptVal to) { Move m; m.kind = Change; m.desc = desc; m.a = e; m.to = to; return m; }
  static Move swap(size_t desc, size_t l, size_t r) { Move m; m.kind = Swap; m.desc = desc; m.a = l; m.b = r; return m; }
  static Move compound(std::vector<ScalarEdit> e) { Move m; m.kind = Compound; m.edits = std::move(e); return m; }
  static Move list_change(size_t desc, size_t se, size_t sp, size_t de, size_t dp) {
    Move m; m.kind = ListChange; m.desc = desc; m.a = se; m.b = sp; m.c = de; m.d = dp; return m;
  }
  static Move list_swap(size_t desc, size_t e1, size_t p1, size_t e2, size_t p2) {
    Move m; m.kind = ListSwap; m.desc = desc; m.a = e1; m.b = p1; m.c = e2; m.d = p2; return m;
  }
  static Move list_reverse(size_t desc, size_t entity, size_t start, size_t end) {  // reverses [start, end)
    Move m; m.kind = ListReverse; m.desc = desc; m.a = entity; m.b = start; m.c = end; return m;
  }
};

// relocates the contiguous segment [start, end) of src_e to dst_e at dst_pos; for an intra-list move dst_pos is
// a position of the list AFTER the removal (sublist_change.rs:17-49)
inline Move move_sublist_change(size_t desc, size_t src_e, size_t start, size_t end, size_t dst_e, size_t dst_pos) {
  Move m; m.kind = Move::SublistChange; m.desc = desc; m.a = src_e; m.b = start; m.c = end; m.d = dst_e; m.e = dst_pos;
  return m;
}
// segment_layout.rs:49-75: the move that puts the segment back
inline Move sublist_change_inverse(const Move& m) {
  return move_sublist_change(m.desc, m.d, m.e, m.e + (m.c - m.b), m.a, m.b);
}

// exchanges the segments [s1, e1) of first_e and [s2, e2) of second_e (sublist_swap.rs:17-42)
inline Move move_sublist_swap(size_t desc, size_t first_e, size_t s1, size_t e1, size_t second_e, size_t s2, size_t e2) {
  Move m; m.kind = Move::SublistSwap; m.desc = desc; m.a = first_e; m.b = s1; m.c = e1; m.d = second_e; m.e = s2; m.f = e2;
  return m;
}
// segment_layout.rs:118-165: where the two segments sit after the swap
inline Move sublist_swap_inverse(const Move& m) {
  const size_t l1 = m.c - m.b, l2 = m.f - m.e;
  if (m.a != m.d) return move_sublist_swap(m.desc, m.a, m.b, m.b + l2, m.d, m.e, m.e + l1);
  if (m.b < m.e) return move_sublist_swap(m.desc, m.a, m.b, m.b + l2, m.d, m.e - l1 + l2, m.e - l1 + l2 + l1);
  return move_sublist_swap(m.desc, m.a, m.b - l2 + l1, m.b - l2 + l1 + l2, m.d, m.e, m.e + l1);
}

// k_opt_reconnection.rs:63-205
struct KOptReconnection {
  uint8_t order[6];
  uint8_t reverse_mask, len;
  size_t k() const { return (size_t)len - 1; }
  bool should_reverse(size_t i) const { return (reverse_mask >> i) & 1; }
  bool is_identity() const {
    if (reverse_mask) return false;
    for (size_t i = 0; i < len; ++i)
      if (order[i] != i) return false;
    return true;
  }
};
// k_opt_reconnection.rs:218-262: middle segments in every order (first item first, recursively), every reversal mask
inline std::vector<KOptReconnection> enumerate_reconnections(size_t k) {
  std::vector<KOptReconnection> out;
  std::vector<uint8_t> middle;
  for (size_t i = 1; i < k; ++i) middle.push_back((uint8_t)i);
  std::vector<std::vector<uint8_t>> perms;
  std::function<void(std::vector<uint8_t>, std::vector<uint8_t>)> rec = [&](std::vector<uint8_t> head, std::vector<uint8_t> rest) {
    if (rest.empty()) {
      perms.push_back(head);
      return;
    }
    for (size_t i = 0; i < rest.size(); ++i) {
      auto h = head;
      h.push_back(rest[i]);
      auto r = rest;
      r.erase(r.begin() + (ptrdiff_t)i);
      rec(h, r);
    }
  };
  rec({}, middle);
  for (auto& perm : perms) {
    KOptReconnection r{};
    r.len = (uint8_t)(k + 1);
    r.order[0] = 0;
    for (size_t i = 0; i < perm.size(); ++i) r.order[i + 1] = perm[i];
    r.order[k] = (uint8_t)k;
    for (uint32_t mask = 0; mask < (1u << (k - 1)); ++mask) {
      r.reverse_mask = (uint8_t)(mask << 1);
      if (!r.is_identity()) out.push_back(r);
    }
  }
  return out;
}
inline Move move_k_opt(size_t desc, size_t entity, std::vector<size_t> cuts, const KOptReconnection& r) {
  Move m;
  m.kind = Move::KOpt;
  m.desc = desc;
  m.a = entity;
  m.cuts = std::move(cuts);
  for (size_t i = 0; i < 6; ++i) m.seg_order[i] = r.order[i];
  m.reverse_mask = r.reverse_mask;
  m.seg_len = r.len;
  return m;
}

struct Undo {
  OptVal v0, v1;
  std::vector<OptVal> many;
  std::vector<size_t> list;  // KOpt: the whole route before the move (k_opt.rs:40-85 returns it as the undo)
};

inline size_t adjusted_destination(const Move& m) {  // list_kernel/change.rs:28-34
  return (m.a == m.c && m.d > m.b) ? m.d - 1 : m.d;
}

template <class S, class Sc>
bool is_doable(const Move& m, ScoreDirector<S, Sc>& dir) {
  S& s = dir.working;
  auto& ac = dir.access;
  switch (m.kind) {
    case Move::Change: {
      OptVal cur = ac.get(s, m.desc, m.a);
      if (!cur && !m.to) return false;
      if (cur && m.to) return *cur != *m.to;
      return true;
    }
    case Move::Swap: return ac.get(s, m.desc, m.a) != ac.get(s, m.desc, m.b);
    case Move::Compound: {
      if (m.edits.empty()) return false;
      bool changes = false;
      for (auto& e : m.edits) changes |= ac.get(s, e.descriptor_index, e.entity_index) != e.to_value;
      return changes;
    }
    case Move::ListChange: {
      size_t src_len = ac.list(s, m.desc, m.a).size();
      if (m.b >= src_len) return false;
      size_t dst_len = ac.list(s, m.desc, m.c).size();
      size_t max_dst = m.a == m.c ? src_len : dst_len;
      if (m.d > max_dst) return false;
      return m.a != m.c || (m.b != m.d && m.d != m.b + 1);
    }
    case Move::ListSwap: {
      auto& l1 = ac.list(s, m.desc, m.a);
      auto& l2 = ac.list(s, m.desc, m.c);
      if (m.b >= l1.size() || m.d >= l2.size() || (m.a == m.c && m.b == m.d)) return false;
      return l1[m.b] != l2[m.d];
    }
    case Move::ListReverse:  // reverse.rs:21-33
      return m.c > m.b + 1 && m.c <= ac.list(s, m.desc, m.a).size();
    case Move::SublistChange: {  // sublist_change.rs:17-49
      if (m.b >= m.c) return false;
      size_t src_len = ac.list(s, m.desc, m.a).size();
      if (m.c > src_len) return false;
      size_t dst_len = ac.list(s, m.desc, m.d).size();
      size_t max_dst = m.a == m.d ? src_len - (m.c - m.b) : dst_len;
      if (m.e > max_dst) return false;
      return m.a != m.d || m.e != m.b;
    }
    case Move::KOpt: {  // k_opt.rs:11-38 (all cuts on one entity)
      const size_t k = m.cuts.size();
      if (k < 2 || (size_t)m.seg_len != k + 1) return false;
      const size_t len = ac.list(s, m.desc, m.a).size();
      for (size_t c : m.cuts)
        if (c > len) return false;
      for (size_t i = 1; i < k; ++i)
        if (m.cuts[i] <= m.cuts[i - 1]) return false;
      return true;
    }
    case Move::SublistSwap: {  // sublist_swap.rs:17-42
      if (m.b >= m.c || m.e >= m.f) return false;
      if (m.c > ac.list(s, m.desc, m.a).size() || m.f > ac.list(s, m.desc, m.d).size()) return false;
      return !(m.a == m.d && m.b < m.f && m.e < m.c);
    }
  }
  return false;
}

inline std::vector<std::pair<size_t, size_t>> unique_affected(const std::vector<ScalarEdit>& edits) {
  std::vector<std::pair<size_t, size_t>> out;
  for (auto& e : edits) {
    auto p = std::make_pair(e.descriptor_index, e.entity_index);
    if (std::find(out.begin(), out.end(), p) == out.end()) out.push_back(p);
  }
  return out;
}

template <class S, class Sc>
Undo do_move(const Move& m, ScoreDirector<S, Sc>& dir) {
  S& s = dir.working;
  auto& ac = dir.access;
  Undo u;
  switch (m.kind) {
    case Move::Change:
      u.v0 = ac.get(s, m.desc, m.a);
      dir.before_variable_changed(m.desc, m.a);
      ac.set(s, m.desc, m.a, m.to);
      dir.after_variable_changed(m.desc, m.a);
      break;
    case Move::Swap:
      u.v0 = ac.get(s, m.desc, m.a);
      u.v1 = ac.get(s, m.desc, m.b);
      dir.before_variable_changed(m.desc, m.a);
      dir.before_variable_changed(m.desc, m.b);
      ac.set(s, m.desc, m.a, u.v1);
      ac.set(s, m.desc, m.b, u.v0);
      dir.after_variable_changed(m.desc, m.a);
      dir.after_variable_changed(m.desc, m.b);
      break;
    case Move::Compound: {
      auto affected = unique_affected(m.edits);
      for (auto& e : m.edits) u.many.push_back(ac.get(s, e.descriptor_index, e.entity_index));
      for (auto& p : affected) dir.before_variable_changed(p.first, p.second);
      for (auto& e : m.edits) ac.set(s, e.descriptor_index, e.entity_index, e.to_value);
      for (auto it = affected.rbegin(); it != affected.rend(); ++it) dir.after_variable_changed(it->first, it->second);
      break;
    }
    case Move::ListChange: {
      bool intra = m.a == m.c;
      dir.before_variable_changed(m.desc, m.a);
      if (!intra) dir.before_variable_changed(m.desc, m.c);
      auto& src = ac.list(s, m.desc, m.a);
      size_t value = src[m.b];
      src.erase(src.begin() + (ptrdiff_t)m.b);
      auto& dst = ac.list(s, m.desc, m.c);
      dst.insert(dst.begin() + (ptrdiff_t)adjusted_destination(m), value);
      dir.after_variable_changed(m.desc, m.a);
      if (!intra) dir.after_variable_changed(m.desc, m.c);
      break;
    }
    case Move::ListSwap: {
      bool intra = m.a == m.c;
      size_t v1 = ac.list(s, m.desc, m.a)[m.b];
      size_t v2 = ac.list(s, m.desc, m.c)[m.d];
      dir.before_variable_changed(m.desc, m.a);
      if (!intra) dir.before_variable_changed(m.desc, m.c);
      ac.list(s, m.desc, m.a)[m.b] = v2;
      ac.list(s, m.desc, m.c)[m.d] = v1;
      dir.after_variable_changed(m.desc, m.a);
      if (!intra) dir.after_variable_changed(m.desc, m.c);
      break;
    }
    case Move::ListReverse: {  // reverse.rs:35-58
      dir.before_variable_changed(m.desc, m.a);
      auto& l = ac.list(s, m.desc, m.a);
      std::reverse(l.begin() + m.b, l.begin() + m.c);
      dir.after_variable_changed(m.desc, m.a);
      break;
    }
    case Move::SublistChange: {  // sublist_change.rs:87-125: source notified first, then destination
      bool intra = m.a == m.d;
      dir.before_variable_changed(m.desc, m.a);
      if (!intra) dir.before_variable_changed(m.desc, m.d);
      auto& src = ac.list(s, m.desc, m.a);
      std::vector<size_t> seg(src.begin() + (ptrdiff_t)m.b, src.begin() + (ptrdiff_t)m.c);
      src.erase(src.begin() + (ptrdiff_t)m.b, src.begin() + (ptrdiff_t)m.c);
      auto& dst = ac.list(s, m.desc, m.d);
      dst.insert(dst.begin() + (ptrdiff_t)m.e, seg.begin(), seg.end());
      dir.after_variable_changed(m.desc, m.a);
      if (!intra) dir.after_variable_changed(m.desc, m.d);
      break;
    }
    case Move::KOpt: {  // k_opt.rs:40-85: cut, re-order, reverse, rebuild the whole route
      dir.before_variable_changed(m.desc, m.a);
      auto& l = ac.list(s, m.desc, m.a);
      u.list = l;
      std::vector<size_t> bounds{0};
      bounds.insert(bounds.end(), m.cuts.begin(), m.cuts.end());
      bounds.push_back(l.size());
      std::vector<size_t> out;
      for (size_t pos = 0; pos < m.seg_len; ++pos) {
        const size_t seg = m.seg_order[pos];
        std::vector<size_t> part(u.list.begin() + (ptrdiff_t)bounds[seg], u.list.begin() + (ptrdiff_t)bounds[seg + 1]);
        if ((m.reverse_mask >> seg) & 1) std::reverse(part.begin(), part.end());
        out.insert(out.end(), part.begin(), part.end());
      }
      l = out;
      dir.after_variable_changed(m.desc, m.a);
      break;
    }
    case Move::SublistSwap: {  // sublist_swap.rs:87-170
      bool intra = m.a == m.d;
      dir.before_variable_changed(m.desc, m.a);
      if (!intra) dir.before_variable_changed(m.desc, m.d);
      if (intra) {  // later segment out first, then the earlier; later elements go to the earlier start
        auto& l = ac.list(s, m.desc, m.a);
        const bool first_early = m.b <= m.e;
        const size_t es = first_early ? m.b : m.e, ee = first_early ? m.c : m.f;
        const size_t ls = first_early ? m.e : m.b, le = first_early ? m.f : m.c;
        std::vector<size_t> late(l.begin() + (ptrdiff_t)ls, l.begin() + (ptrdiff_t)le);
        l.erase(l.begin() + (ptrdiff_t)ls, l.begin() + (ptrdiff_t)le);
        std::vector<size_t> early(l.begin() + (ptrdiff_t)es, l.begin() + (ptrdiff_t)ee);
        l.erase(l.begin() + (ptrdiff_t)es, l.begin() + (ptrdiff_t)ee);
        l.insert(l.begin() + (ptrdiff_t)es, late.begin(), late.end());
        const size_t new_late = ls - early.size() + late.size();
        l.insert(l.begin() + (ptrdiff_t)new_late, early.begin(), early.end());
      } else {
        auto& l1 = ac.list(s, m.desc, m.a);
        auto& l2 = ac.list(s, m.desc, m.d);
        std::vector<size_t> s1(l1.begin() + (ptrdiff_t)m.b, l1.begin() + (ptrdiff_t)m.c);
        l1.erase(l1.begin() + (ptrdiff_t)m.b, l1.begin() + (ptrdiff_t)m.c);
        std::vector<size_t> s2(l2.begin() + (ptrdiff_t)m.e, l2.begin() + (ptrdiff_t)m.f);
        l2.erase(l2.begin() + (ptrdiff_t)m.e, l2.begin() + (ptrdiff_t)m.f);
        l1.insert(l1.begin() + (ptrdiff_t)m.b, s2.begin(), s2.end());
        l2.insert(l2.begin() + (ptrdiff_t)m.e, s1.begin(), s1.end());
      }
      dir.after_variable_changed(m.desc, m.a);
      if (!intra) dir.after_variable_changed(m.desc, m.d);
      break;
    }
  }
  return u;
}

template <class S, class Sc>
void undo_move(const Move& m, ScoreDirector<S, Sc>& dir, const Undo& u) {
  S& s = dir.working;
  auto& ac = dir.access;
  switch (m.kind) {
    case Move::Change:
      dir.before_variable_changed(m.desc, m.a);
      ac.set(s, m.desc, m.a, u.v0);
      dir.after_variable_changed(m.desc, m.a);
      break;
    case Move::Swap:
      dir.before_variable_changed(m.desc, m.a);
      dir.before_variable_changed(m.desc, m.b);
      ac.set(s, m.desc, m.a, u.v0);
      ac.set(s, m.desc, m.b, u.v1);
      dir.after_variable_changed(m.desc, m.a);
      dir.after_variable_changed(m.desc, m.b);
      break;
    case Move::Compound: {
      auto affected = unique_affected(m.edits);
      for (auto& p : affected) dir.before_variable_changed(p.first, p.second);
      for (size_t i = 0; i < m.edits.size(); ++i)
        ac.set(s, m.edits[i].descriptor_index, m.edits[i].entity_index, u.many[i]);
      for (auto it = affected.rbegin(); it != affected.rend(); ++it) dir.after_variable_changed(it->first, it->second);
      break;
    }
    case Move::ListChange: {  // list_kernel/change.rs:122-153 — destination notified first
      bool intra = m.a == m.c;
      dir.before_variable_changed(m.desc, m.c);
      if (!intra) dir.before_variable_changed(m.desc, m.a);
      auto& dst = ac.list(s, m.desc, m.c);
      size_t adj = adjusted_destination(m);
      size_t value = dst[adj];
      dst.erase(dst.begin() + (ptrdiff_t)adj);
      auto& src = ac.list(s, m.desc, m.a);
      src.insert(src.begin() + (ptrdiff_t)m.b, value);
      dir.after_variable_changed(m.desc, m.c);
      if (!intra) dir.after_variable_changed(m.desc, m.a);
      break;
    }
    case Move::ListSwap:     // a swap is its own inverse
    case Move::ListReverse: {  // so is a reversal
      do_move(m, dir);
      break;
    }
    case Move::SublistChange: {  // sublist_change.rs:69-85: the inverse layout applied the same way
      do_move(sublist_change_inverse(m), dir);
      break;
    }
    case Move::SublistSwap: {  // sublist_swap.rs:66-85
      do_move(sublist_swap_inverse(m), dir);
      break;
    }
    case Move::KOpt: {  // k_opt.rs:87-110: the saved route goes back
      dir.before_variable_changed(m.desc, m.a);
      ac.list(s, m.desc, m.a) = u.list;
      dir.after_variable_changed(m.desc, m.a);
      break;
    }
  }
}

// evaluation.rs:20-115
enum class EvalKind { Scored, NotDoable, RejectedByHardImprovement, RejectedByScoreImprovement };
template <class Sc>
struct CandidateEvaluation {
  EvalKind kind;
  Sc score;
};

template <class S, class Sc>
CandidateEvaluation<Sc> evaluate_candidate(const Move& m, ScoreDirector<S, Sc>& dir, Sc reference_score) {
  if (!is_doable(m, dir)) return {EvalKind::NotDoable, Sc::zero()};
  auto state = dir.snapshot_score_state();
  Undo u = do_move(m, dir);
  Sc move_score = dir.calculate_score();
  undo_move(m, dir, u);
  dir.restore_score_state(state);
  dir.score_calculations += 1;
  HardDelta hd = hard_score_delta(reference_score, move_score);
  if (m.requires_hard_improvement && hd != HardDelta::Improving) return {EvalKind::RejectedByHardImprovement, move_score};
  if (m.requires_score_improvement && move_score <= reference_score) return {EvalKind::RejectedByScoreImprovement, move_score};
  return {EvalKind::Scored, move_score};
}

// ---------------------------------------------------------------------------------------------
inline uint64_t splitmix64(uint64_t v) {  // iter.rs:193-198
  v += 0x9E3779B97F4A7C15ull;
  v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
  v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
  return v ^ (v >> 31);
}
inline size_t gcd_sz(size_t l, size_t r) {
  while (r != 0) {
    size_t t = l % r;
    l = r;
    r = t;
  }
  return l;
}

enum class SelectionOrder { Original, Sorted, Probabilistic, Random, Shuffled };

struct MoveStreamContext {  // iter.rs:14-184
  uint64_t step_index = 0, step_seed = 0;
  SelectionOrder order = SelectionOrder::Original;
  bool is_canonical() const {
    return order == SelectionOrder::Original || order == SelectionOrder::Sorted || order == SelectionOrder::Probabilistic;
  }
  uint64_t mixed_seed(uint64_t salt) const {
    return splitmix64(step_seed ^ (step_index * 0x9E3779B97F4A7C15ull) ^ salt);
  }
  size_t random_index(size_t len, uint64_t salt) const { return len <= 1 ? 0 : (size_t)(mixed_seed(salt) % len); }
  size_t random_stride(size_t len, uint64_t salt) const {
    if (len <= 1) return 1;
    size_t stride = (size_t)(mixed_seed(salt) % (len - 1)) + 1;
    while (gcd_sz(stride, len) != 1) stride = stride == len - 1 ? 1 : stride + 1;
    return stride;
  }
  size_t selection_index(size_t offset, size_t len, uint64_t salt) const {
    switch (order) {
      case SelectionOrder::Random:
        return random_index(len, salt ^ ((uint64_t)offset * 0xD1B54A32D192ED03ull));
      case SelectionOrder::Shuffled: {
        size_t start = random_index(len, salt);
        size_t stride = random_stride(len, salt ^ 0xA24BAED4963EE407ull);
        return (start + offset * stride) % len;
      }
      default: return offset;
    }
  }
  // iter.rs:130-147: an outer source dimension is walked once — Random and Shuffled share the strided permutation
  size_t selection_index_without_replacement(size_t offset, size_t len, uint64_t salt) const {
    if (is_canonical()) return offset;
    size_t start = random_index(len, salt);
    size_t stride = random_stride(len, salt ^ 0xA24BAED4963EE407ull);
    return (start + offset * stride) % len;
  }
};

// move_selector/change.rs:66-104,246-307 — values in order, then the to-None move when
// allows_unassigned and the entity is currently assigned.
template <class S>
std::vector<Move> enumerate_change_moves(const S& s, const Access<S>& ac, size_t desc, size_t variable_index,
                                         size_t n_values, bool allows_unassigned, MoveStreamContext ctx) {
  std::vector<Move> out;
  size_t n = ac.entity_count(s, desc);
  uint64_t entity_salt = 0xC4A46E0000000001ull ^ ((uint64_t)desc << 32) ^ (uint64_t)variable_index;
  for (size_t eo = 0; eo < n; ++eo) {
    size_t e = ctx.selection_index(eo, n, entity_salt);
    bool assigned = ac.get(s, desc, e).has_value();
    uint64_t value_salt = 0xC4A46E0000000000ull ^ (uint64_t)e ^ ((uint64_t)desc << 32) ^ (uint64_t)variable_index;
    for (size_t vo = 0; vo < n_values; ++vo) {
      size_t v = ctx.is_canonical() ? vo : ctx.selection_index(vo, n_values, value_salt);
      out.push_back(Move::change(desc, e, OptVal(v)));
    }
    if (allows_unassigned && assigned) out.push_back(Move::change(desc, e, std::nullopt));
  }
  return out;
}

// heuristic/selector/move_selector/swap.rs:64-100,196-233 (SwapMoveSelector over one entity class on both sides):
// the left and the right entity lists are permuted independently by the stream context; every (left, right)
// pair with left.entity_index < right.entity_index is a SwapMove, left-major.
template <class S>
std::vector<Move> enumerate_swap_moves(const S& s, const Access<S>& ac, size_t desc, size_t variable_index,
                                       MoveStreamContext ctx) {
  std::vector<Move> out;
  const size_t n = ac.entity_count(s, desc);
  const uint64_t salt = ((uint64_t)desc << 32) ^ (uint64_t)variable_index;
  std::vector<size_t> left(n), right(n);
  for (size_t o = 0; o < n; ++o) {
    left[o] = ctx.selection_index(o, n, 0x5A09000000000001ull ^ salt);
    right[o] = ctx.selection_index(o, n, 0x5A09000000000002ull ^ salt);
  }
  for (size_t lo = 0; lo < n; ++lo)
    for (size_t ro = 0; ro < n; ++ro)
      if (left[lo] < right[ro]) out.push_back(Move::swap(desc, left[lo], right[ro]));
  return out;
}

// nearby_list_support.rs:3-34 — stable bounded insertion sort on the f64 distance.
struct NearbyCandidate {
  size_t entity, position;
  double distance;
};
inline void sort_and_limit_nearby_candidates(std::vector<NearbyCandidate>& c, size_t max_nearby) {
  if (max_nearby == 0) {
    c.clear();
    return;
  }
  size_t retained = 0;
  for (size_t read = 0; read < c.size(); ++read) {
    NearbyCandidate cand = c[read];
    // partition_point(existing.distance <= cand.distance), NaN compares Equal.
    size_t lo = 0, hi = retained;
    while (lo < hi) {
      size_t mid = lo + (hi - lo) / 2;
      bool greater = c[mid].distance > cand.distance;
      if (!greater) lo = mid + 1; else hi = mid;
    }
    size_t insertion = lo;
    if (insertion >= max_nearby) continue;
    size_t next_retained = std::min(retained + 1, max_nearby);
    for (size_t i = next_retained - 1; i > insertion; --i) c[i] = c[i - 1];
    c[insertion] = cand;
    retained = next_retained;
  }
  c.resize(retained);
}

// list_kernel/nearby_change.rs:102-232 (no owner restriction, no precedence graph).
// distance(src_e, src_p, dst_e, ref_p) is the CrossEntityDistanceMeter.
template <class S, class Dist>
std::vector<Move> enumerate_nearby_list_change_moves(S& s, const Access<S>& ac, size_t desc, size_t max_nearby,
                                                     MoveStreamContext ctx, Dist distance) {
  constexpr uint64_t ENTITY_SALT = 0xA1EA2B17C4A40001ull, SOURCE_SALT = 0xA1EA2B17C4A40002ull;
  size_t n = ac.entity_count(s, desc);
  std::vector<size_t> entities(n), lens(n);
  for (size_t o = 0; o < n; ++o) {
    size_t e = n <= 1 ? o : ctx.selection_index(o, n, ENTITY_SALT ^ (uint64_t)desc);
    entities[o] = e;
    lens[o] = ac.list(s, desc, e).size();
  }
  std::vector<Move> out;
  std::vector<NearbyCandidate> cand;
  for (size_t si = 0; si < n; ++si) {
    size_t se = entities[si], slen = lens[si];
    for (size_t po = 0; po < slen; ++po) {
      size_t sp = ctx.selection_index(po, slen, SOURCE_SALT ^ (uint64_t)se ^ (uint64_t)desc);
      cand.clear();
      for (size_t dp = 0; dp <= slen; ++dp) {
        if (dp == sp || dp == sp + 1) continue;
        double dist = distance(s, se, sp, se, std::min(dp, slen > 0 ? slen - 1 : 0));
        if (std::isfinite(dist)) cand.push_back({se, dp, dist});
      }
      for (size_t di = 0; di < n; ++di) {
        if (di == si) continue;
        size_t de = entities[di], dlen = lens[di];
        for (size_t dp = 0; dp <= dlen; ++dp) {
          double dist = distance(s, se, sp, de, std::min(dp, dlen > 0 ? dlen - 1 : 0));
          if (std::isfinite(dist)) cand.push_back({de, dp, dist});
        }
      }
      sort_and_limit_nearby_candidates(cand, max_nearby);
      for (auto& c : cand) out.push_back(Move::list_change(desc, se, sp, c.entity, c.position));
    }
  }
  return out;
}

// list_kernel/nearby_swap.rs:99-262 (NearbySwapCursor; no owner restriction, no precedence graph,
// not fixed to the current entity): for every source (entity order, position order from the stream
// context) the destinations are the later positions of the same list and every position of the
// entities later in the entity ORDER, ranked by the meter with the stable bounded top-k; sources
// without destinations are skipped.
template <class S, class Dist>
std::vector<Move> enumerate_nearby_list_swap_moves(S& s, const Access<S>& ac, size_t desc, size_t max_nearby,
                                                   MoveStreamContext ctx, Dist distance) {
  constexpr uint64_t ENTITY_SALT = 0xA1EA25A090000001ull, SOURCE_SALT = 0xA1EA25A090000002ull;
  size_t n = ac.entity_count(s, desc);
  std::vector<size_t> entities(n), lens(n);
  for (size_t o = 0; o < n; ++o) {
    size_t e = n <= 1 ? o : ctx.selection_index(o, n, ENTITY_SALT ^ (uint64_t)desc);
    entities[o] = e;
    lens[o] = ac.list(s, desc, e).size();
  }
  std::vector<Move> out;
  std::vector<NearbyCandidate> cand;
  for (size_t si = 0; si < n; ++si) {
    size_t se = entities[si], slen = lens[si];
    for (size_t po = 0; po < slen; ++po) {
      size_t sp = ctx.selection_index(po, slen, SOURCE_SALT ^ (uint64_t)se ^ (uint64_t)desc);
      cand.clear();
      for (size_t dp = sp + 1; dp < slen; ++dp) {
        double dist = distance(s, se, sp, se, dp);
        if (std::isfinite(dist)) cand.push_back({se, dp, dist});
      }
      for (size_t di = si + 1; di < n; ++di) {
        size_t de = entities[di], dlen = lens[di];
        for (size_t dp = 0; dp < dlen; ++dp) {
          double dist = distance(s, se, sp, de, dp);
          if (std::isfinite(dist)) cand.push_back({de, dp, dist});
        }
      }
      sort_and_limit_nearby_candidates(cand, max_nearby);
      for (auto& c : cand) out.push_back(Move::list_swap(desc, se, sp, c.entity, c.position));
    }
  }
  return out;
}

// heuristic/selector/list_reverse.rs:139-172 + list_kernel/reverse.rs:66-108 (ReverseCursor): entities in
// stream order; per entity with len >= 2 every start (stream order) and every end in start+2..=len
// (stream order over the end_count = len - start - 1 choices).
template <class S>
std::vector<Move> enumerate_list_reverse_moves(S& s, const Access<S>& ac, size_t desc, MoveStreamContext ctx) {
  constexpr uint64_t ENTITY_SALT = 0x11572A0700000001ull, START_SALT = 0x11572A0700000002ull,
                     END_SALT = 0x11572A0700000003ull;
  size_t n = ac.entity_count(s, desc);
  std::vector<Move> out;
  for (size_t o = 0; o < n; ++o) {
    size_t e = ctx.selection_index(o, n, ENTITY_SALT ^ (uint64_t)desc);
    size_t len = ac.list(s, desc, e).size();
    if (len < 2) continue;
    for (size_t so = 0; so < len; ++so) {
      size_t start = ctx.selection_index(so, len, START_SALT ^ (uint64_t)e ^ (uint64_t)desc);
      size_t end_count = len > start + 1 ? len - (start + 1) : 0;
      for (size_t eo = 0; eo < end_count; ++eo) {
        size_t end = start + 2 + ctx.selection_index(eo, end_count, END_SALT ^ (uint64_t)e ^ (uint64_t)start);
        out.push_back(Move::list_reverse(desc, e, start, end));
      }
    }
  }
  return out;
}

// heuristic/selector/k_opt/iterators.rs:93-176: binomial, number of cut combinations, and the combination of a given
// rank (lexicographic over the k chosen slots; position = choice + min_seg + index * (min_seg - 1))
inline size_t binomial(size_t n, size_t k) {
  if (k > n) return 0;
  if (k == 0 || k == n) return 1;
  k = std::min(k, n - k);
  size_t r = 1;
  for (size_t i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
  return r;
}
inline size_t count_cut_combinations(size_t k, size_t len, size_t min_seg) {
  const size_t min_len = (k + 1) * min_seg;
  return len < min_len ? 0 : binomial(len - min_len + k, k);
}
inline bool cut_combination_at(size_t k, size_t len, size_t min_seg, size_t rank, std::vector<size_t>& cuts) {
  cuts.clear();
  if (k == 0 || min_seg == 0 || len < (k + 1) * min_seg) return false;
  const size_t choice_count = len - (k + 1) * min_seg + k;
  if (rank >= binomial(choice_count, k)) return false;
  size_t start = 0;
  for (size_t position = 0; position < k; ++position) {
    const size_t remaining = k - position - 1, maximum = choice_count - (k - position);
    bool found = false;
    for (size_t cand = start; cand <= maximum; ++cand) {
      const size_t suffix = binomial(choice_count - cand - 1, remaining);
      if (rank < suffix) {
        cuts.push_back(cand + min_seg + position * (min_seg - 1));
        start = cand + 1;
        found = true;
        break;
      }
      rank -= suffix;
    }
    if (!found) return false;
  }
  return true;
}
// heuristic/selector/list_kernel/k_opt/full.rs:34-98 (KOptCursor): per entity the moves are
// (cut combination rank) x (reconnection pattern), pulled through selection_index over their product; 3-opt uses
// the static THREE_OPT_RECONNECTIONS table (selector.rs:111-115), which enumerate_reconnections(3) reproduces.
// Entities in the without-replacement stream order (full.rs:48-51, salt 0x4B0F7E1171000001 ^ descriptor).
template <class S>
std::vector<Move> enumerate_k_opt_moves(S& s, const Access<S>& ac, size_t desc, size_t k, size_t min_seg,
                                        MoveStreamContext ctx) {
  const auto patterns = enumerate_reconnections(k);
  std::vector<Move> out;
  std::vector<size_t> cuts;
  const size_t n = ac.entity_count(s, desc);
  for (size_t eo = 0; eo < n; ++eo) {
    const size_t e = ctx.selection_index_without_replacement(eo, n, 0x4B0F7E1171000001ull ^ (uint64_t)desc);
    const size_t len = ac.list(s, desc, e).size();
    const size_t move_count = count_cut_combinations(k, len, min_seg) * patterns.size();
    for (size_t off = 0; off < move_count; ++off) {
      const size_t sel = ctx.selection_index(off, move_count, 0x4B0F7E1171000002ull ^ (uint64_t)desc ^ (uint64_t)e);
      if (!cut_combination_at(k, len, min_seg, sel / patterns.size(), cuts)) throw std::logic_error("k-opt cut rank");
      out.push_back(move_k_opt(desc, e, cuts, patterns[sel % patterns.size()]));
      out.back().f = sel % patterns.size();  // pattern index inside enumerate_reconnections(k)
    }
  }
  return out;
}

// heuristic/selector/sublist_swap.rs:160-190 + list_kernel/sublist_swap.rs:28-318 (SublistSwapCursor, no owner
// restriction, no precedence graph): first segments in entity / start / size stream order; second segments from
// the same entity onwards (entity order); inside one list only segments starting at or after the first one's end.
template <class S>
std::vector<Move> enumerate_sublist_swap_moves(S& s, const Access<S>& ac, size_t desc, size_t min_size, size_t max_size,
                                               MoveStreamContext ctx) {
  constexpr uint64_t ENTITY_SALT = 0x5B1575A090000001ull, START_SALT = 0x5B1575A090000002ull,
                     SIZE_SALT = 0x5B1575A090000003ull;
  size_t n = ac.entity_count(s, desc);
  std::vector<size_t> entities(n), lens(n);
  for (size_t o = 0; o < n; ++o) {
    size_t e = n <= 1 ? o : ctx.selection_index(o, n, ENTITY_SALT ^ (uint64_t)desc);
    entities[o] = e;
    lens[o] = ac.list(s, desc, e).size();
  }
  auto segments = [&](size_t idx) {  // SublistSegmentCursor
    std::vector<std::pair<size_t, size_t>> out;
    const size_t e = entities[idx], len = lens[idx];
    if (len < min_size) return out;
    for (size_t so = 0; so < len; ++so) {
      const size_t start = ctx.selection_index(so, len, START_SALT ^ (uint64_t)e ^ (uint64_t)desc);
      const size_t max_valid = std::min(max_size, len - start);
      if (max_valid < min_size) continue;
      const size_t count = max_valid - min_size + 1;
      for (size_t zo = 0; zo < count; ++zo)
        out.push_back({start, start + min_size + ctx.selection_index(zo, count, SIZE_SALT ^ (uint64_t)e ^ (uint64_t)start)});
    }
    return out;
  };
  std::vector<std::vector<std::pair<size_t, size_t>>> segs(n);
  for (size_t i = 0; i < n; ++i) segs[i] = segments(i);
  std::vector<Move> out;
  for (size_t fi = 0; fi < n; ++fi)
    for (auto& first : segs[fi])
      for (size_t si = fi; si < n; ++si)
        for (auto& second : segs[si]) {
          if (fi == si && (second.first < first.second || (first == second))) continue;
          out.push_back(move_sublist_swap(desc, entities[fi], first.first, first.second, entities[si], second.first, second.second));
        }
  return out;
}

// heuristic/selector/sublist_change.rs:166-205 + list_kernel/sublist_change.rs:103-268 (SublistChangeCursor, no
// owner restriction, no precedence graph): entities in stream order; per source every segment start (stream
// order), every valid size min..=max (stream order); intra-list destinations over the post-removal list
// (skipping the segment's own start) before the insertions into every other entity in entity order.
template <class S>
std::vector<Move> enumerate_sublist_change_moves(S& s, const Access<S>& ac, size_t desc, size_t min_size,
                                                 size_t max_size, MoveStreamContext ctx) {
  constexpr uint64_t ENTITY_SALT = 0x5B157C4A46E00001ull, START_SALT = 0x5B157C4A46E00002ull,
                     SIZE_SALT = 0x5B157C4A46E00003ull, INTRA_SALT = 0x5B157C4A46E00004ull,
                     INTER_SALT = 0x5B157C4A46E00005ull;
  size_t n = ac.entity_count(s, desc);
  std::vector<size_t> entities(n), lens(n);
  for (size_t o = 0; o < n; ++o) {
    size_t e = n <= 1 ? o : ctx.selection_index(o, n, ENTITY_SALT ^ (uint64_t)desc);
    entities[o] = e;
    lens[o] = ac.list(s, desc, e).size();
  }
  std::vector<Move> out;
  for (size_t si = 0; si < n; ++si) {
    const size_t se = entities[si], slen = lens[si];
    if (slen < min_size) continue;
    for (size_t so = 0; so < slen; ++so) {
      const size_t start = ctx.selection_index(so, slen, START_SALT ^ (uint64_t)se ^ (uint64_t)desc);
      const size_t max_valid = std::min(max_size, slen - start);
      const size_t size_count = max_valid >= min_size ? max_valid - min_size + 1 : 0;
      for (size_t zo = 0; zo < size_count; ++zo) {
        const size_t size = min_size + ctx.selection_index(zo, size_count, SIZE_SALT ^ (uint64_t)se ^ (uint64_t)start);
        const size_t end = start + size, post = slen - size;
        for (size_t po = 0; po <= post; ++po) {
          const size_t dp = ctx.selection_index(po, post + 1, INTRA_SALT ^ (uint64_t)se ^ (uint64_t)start);
          if (dp == start) continue;
          out.push_back(move_sublist_change(desc, se, start, end, se, dp));
        }
        for (size_t di = 0; di < n; ++di) {
          if (di == si) continue;
          const size_t de = entities[di], dlen = lens[di];
          for (size_t po = 0; po <= dlen; ++po) {
            const size_t dp = ctx.selection_index(po, dlen + 1, INTER_SALT ^ (uint64_t)se ^ (uint64_t)de ^ (uint64_t)start);
            out.push_back(move_sublist_change(desc, se, start, end, de, dp));
          }
        }
      }
    }
  }
  return out;
}


// ---------------------------------------------------------------------------------------------
// Union of selector families: heuristic/selector/decorator/vec_union.rs:190-366 (UnionScheduler). Children are
// finite streams here (sizes[c] candidates each); the result is the union's pull order as (child, child-local
// index) pairs — CandidateId of the union cursor = position in this list (vec_union.rs:447-455).
enum class UnionOrder { Sequential, RoundRobin, RotatingRoundRobin, Random, StratifiedRandom };

struct UnionScheduler {
  size_t current_cursor = 0, live = 0, cursor_offset = 0, cursor_stride = 1, count = 0;
  UnionOrder order;
  std::vector<bool> exhausted;
  MoveStreamContext ctx;
  uint64_t random_draw = 0, total_live_weight = 0;
  std::vector<uint64_t> weights;
  std::vector<__int128> weighted_current;

  UnionScheduler(size_t cursor_count, UnionOrder o, MoveStreamContext c, std::vector<uint64_t> w)
      : count(cursor_count), order(o), ctx(c), weights(std::move(w)) {
    if (weights.size() != count) throw std::logic_error("union weight count must match child count");
    for (size_t i = 0; i < count; ++i) {
      exhausted.push_back(weights[i] == 0);
      if (weights[i] != 0) ++live;
      total_live_weight += weights[i];
    }
    if (order == UnionOrder::RotatingRoundRobin || order == UnionOrder::StratifiedRandom)
      cursor_offset = ctx.random_index(count, 0xA11CE5E1EC700001ull);
    if (order == UnionOrder::StratifiedRandom) cursor_stride = ctx.random_stride(count, 0xA11CE5E1EC700002ull);
    current_cursor = order == UnionOrder::StratifiedRandom ? 0 : cursor_offset;
    weighted_current.assign(count, 0);
  }

  // next_child(c) -> true when child c yields another candidate
  template <class F>
  bool next(F next_child, size_t& out_child) {
    switch (order) {
      case UnionOrder::Sequential:
        while (current_cursor < count) {
          if (next_child(current_cursor)) {
            out_child = current_cursor;
            return true;
          }
          ++current_cursor;
        }
        return false;
      case UnionOrder::RoundRobin:
      case UnionOrder::RotatingRoundRobin:
        while (live > 0) {
          const size_t c = current_cursor % count;
          current_cursor = (current_cursor + 1) % count;
          if (exhausted[c]) continue;
          if (next_child(c)) {
            out_child = c;
            return true;
          }
          exhausted[c] = true;
          --live;
        }
        return false;
      case UnionOrder::Random:
        while (live > 0) {
          const uint64_t draw = ctx.mixed_seed(0xA11CE5E1EC701000ull + random_draw) % total_live_weight;
          ++random_draw;
          uint64_t cumulative = 0;
          size_t c = count;
          for (size_t i = 0; i < count; ++i) {
            if (exhausted[i]) continue;
            cumulative += weights[i];
            if (draw < cumulative) {
              c = i;
              break;
            }
          }
          if (c == count) throw std::logic_error("random union live weights must select one child");
          if (next_child(c)) {
            out_child = c;
            return true;
          }
          exhausted[c] = true;
          --live;
          total_live_weight -= weights[c];
        }
        return false;
      case UnionOrder::StratifiedRandom:
        while (live > 0) {
          size_t sel = count;
          __int128 sel_w = 0;
          for (size_t position = 0; position < count; ++position) {
            const size_t c = (cursor_offset + position * cursor_stride) % count;
            if (exhausted[c]) continue;
            weighted_current[c] += (__int128)weights[c];
            if (sel == count || weighted_current[c] > sel_w) {
              sel = c;
              sel_w = weighted_current[c];
            }
          }
          weighted_current[sel] -= (__int128)total_live_weight;
          if (next_child(sel)) {
            out_child = sel;
            return true;
          }
          exhausted[sel] = true;
          --live;
          total_live_weight -= weights[sel];
        }
        return false;
    }
    return false;
  }
};

inline std::vector<std::pair<size_t, size_t>> union_pull_order(const std::vector<size_t>& sizes, UnionOrder order,
                                                               MoveStreamContext ctx, std::vector<uint64_t> weights,
                                                               size_t limit = SIZE_MAX) {
  UnionScheduler sched(sizes.size(), order, ctx, std::move(weights));
  std::vector<size_t> at(sizes.size(), 0);
  std::vector<std::pair<size_t, size_t>> out;
  size_t c = 0;
  while (out.size() < limit && sched.next([&](size_t k) { return at[k] < sizes[k]; }, c)) {
    out.push_back({c, at[c]});
    ++at[c];
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// Foragers (forager.rs). CandidateId = pull index.
inline bool reservoir_pick(uint64_t step_seed, uint64_t equal_count) {  // forager.rs:143-148
  uint64_t mixed = splitmix64(step_seed ^ (equal_count * 0x9E3779B97F4A7C15ull) ^ 0xF04A63E239B74D11ull);
  return mixed % equal_count == 0;
}

template <class Sc>
struct BestCandidate {  // forager.rs:70-141
  bool has = false;
  size_t index = 0;
  Sc score = Sc::zero();
  uint64_t equal_count = 0, step_seed = 0;
  bool random_ties = true;
  void reset(uint64_t seed) {
    has = false;
    equal_count = 0;
    step_seed = seed;
  }
  void consider(size_t idx, Sc sc) {
    if (!has) {
      has = true;
      index = idx;
      score = sc;
      equal_count = 1;
      return;
    }
    if (sc < score) return;
    if (sc > score) {
      index = idx;
      score = sc;
      equal_count = 1;
      return;
    }
    equal_count += 1;
    if (random_ties && reservoir_pick(step_seed, equal_count)) {
      index = idx;
      score = sc;
    }
  }
};

enum class ForagerKind { AcceptedCount, FirstAccepted, BestScore, FirstBestScoreImproving, FirstLastStepScoreImproving };

template <class Sc>
struct Forager {  // forager.rs:167-425, forager/improving.rs:17-227
  ForagerKind kind = ForagerKind::BestScore;
  size_t accepted_count_limit = 1;
  size_t accepted_count = 0;
  BestCandidate<Sc> best;
  Sc best_score = Sc::zero(), last_step_score = Sc::zero();
  bool found_improving = false;
  bool has_improving_limit = false;  // FirstLastStepScoreImproving::with_accepted_count_limit
  void step_started(Sc best_sc, Sc last_sc, uint64_t step_seed) {
    accepted_count = 0;
    best.reset(step_seed);
    best_score = best_sc;
    last_step_score = last_sc;
    found_improving = false;
  }
  void add_move_index(size_t idx, Sc sc) {
    switch (kind) {
      case ForagerKind::AcceptedCount:
        if (accepted_count >= accepted_count_limit) return;
        accepted_count += 1;
        best.consider(idx, sc);
        return;
      case ForagerKind::FirstAccepted:
        if (!best.has) {
          best.has = true;
          best.index = idx;
          best.score = sc;
        }
        return;
      case ForagerKind::BestScore: best.consider(idx, sc); return;
      case ForagerKind::FirstBestScoreImproving:  // improving.rs:86-95
        if (sc > best_score) {
          found_improving = true;
          best.has = true;
          best.index = idx;
          best.score = sc;
          best.equal_count = 1;
          return;
        }
        if (found_improving) return;
        best.consider(idx, sc);
        return;
      case ForagerKind::FirstLastStepScoreImproving:  // improving.rs:197-211
        if (found_improving || (has_improving_limit && accepted_count >= accepted_count_limit)) return;
        accepted_count += 1;
        if (sc > last_step_score) {
          found_improving = true;
          best.has = true;
          best.index = idx;
          best.score = sc;
          best.equal_count = 1;
          return;
        }
        best.consider(idx, sc);
        return;
    }
  }
  bool is_quit_early() const {
    switch (kind) {
      case ForagerKind::AcceptedCount: return accepted_count >= accepted_count_limit;
      case ForagerKind::FirstAccepted: return best.has;
      case ForagerKind::BestScore: return false;
      case ForagerKind::FirstBestScoreImproving: return found_improving;
      default: return found_improving || (has_improving_limit && accepted_count >= accepted_count_limit);
    }
  }
};

// Tabu metadata of a move (heuristic/move/metadata.rs:12-107). The reference hashes variable names and
// Debug-formatted values with SipHash (DefaultHasher) into u64 ids; the ids are only ever compared for
// equality, so this restatement uses the raw indices (injective) and a numeric scope id instead.
struct TabuSignature {
  uint64_t scope = 0;  // (descriptor_index, variable) pair
  std::vector<uint64_t> entity_ids, value_ids;
  std::vector<uint64_t> move_id, undo_move_id;
};
static constexpr uint64_t TABU_NONE_ID = UINT64_MAX;            // metadata.rs:7
static constexpr uint64_t TABU_OP_SWAP = 0xF000000000000001ull;  // metadata.rs:150

// TabuMemory (tabu_search.rs:64-101): FIFO of at most `tenure` entries; tenure 0 = dimension off.
template <class T>
struct TabuMemory {
  size_t tenure = 0;
  std::vector<T> entries;
  bool contains(const T& e) const {
    for (const T& x : entries)
      if (x == e) return true;
    return false;
  }
  void record(const T& e) {
    if (!tenure) return;
    if (entries.size() >= tenure) entries.erase(entries.begin());
    entries.push_back(e);
  }
};

enum class AcceptorKind {
  HillClimbing, LateAcceptance, SimulatedAnnealing, AcceptAll,
  GreatDeluge, StepCountingHillClimbing, DiversifiedLateAcceptance, TabuSearch
};

template <class Sc>
struct Acceptor {
  AcceptorKind kind = AcceptorKind::HillClimbing;
  // LateAcceptance (late_acceptance.rs:89-126) and DiversifiedLateAcceptance share the ring
  std::vector<Sc> history;
  size_t history_idx = 0;
  // SimulatedAnnealing (simulated_annealing.rs): the uniform stream is injected by the caller
  // because the reference draws from rand::SmallRng (third party, unpinned — SURVEY §8c).
  std::function<double()> uniform;
  double decay = 0.999985, hill_climbing_temperature = 1e-9, fallback_temperature = 1.0;
  size_t calibration_samples = 128;
  double target_acceptance = 0.80;
  std::vector<double> level_temperature, sample_sum;
  std::vector<size_t> sample_count;
  bool calibrated = false;
  bool never_accept_hard_regression = false;
  // GreatDeluge (great_deluge.rs:52-101)
  double rain_speed = 0.001;
  bool has_water = false;
  Sc water_level = Sc::zero(), initial_abs_score = Sc::zero();
  // StepCountingHillClimbing (step_counting.rs:55-104)
  uint64_t step_count_limit = 100, steps_since_improvement = 0;
  // DiversifiedLateAcceptance (diversified_late_acceptance.rs:71-140) and TabuSearch aspiration
  double tolerance = 0.01;
  bool has_best = false;
  Sc best_score = Sc::zero();
  // TabuSearch (tabu_search.rs:103-237); a zero tenure disables the dimension (Option::None)
  TabuMemory<std::pair<uint64_t, uint64_t>> entity_memory, value_memory;  // (scope, id) tokens
  TabuMemory<std::vector<uint64_t>> move_memory, reverse_move_memory;
  bool aspiration_enabled = true;

  bool requires_move_signatures() const { return kind == AcceptorKind::TabuSearch; }

  void phase_started(Sc initial, size_t late_size = 400) {
    if (kind == AcceptorKind::LateAcceptance || kind == AcceptorKind::DiversifiedLateAcceptance) {
      history.assign(late_size, initial);
      history_idx = 0;
    }
    if (kind == AcceptorKind::SimulatedAnnealing) {
      level_temperature.assign(Sc::levels, 0.0);
      sample_sum.assign(Sc::levels, 0.0);
      sample_count.assign(Sc::levels, 0);
      calibrated = false;
    }
    if (kind == AcceptorKind::GreatDeluge) {
      has_water = true;
      water_level = initial;
      initial_abs_score = initial.abs();
    }
    if (kind == AcceptorKind::StepCountingHillClimbing) steps_since_improvement = 0;
    if (kind == AcceptorKind::TabuSearch) {
      entity_memory.entries.clear();
      value_memory.entries.clear();
      move_memory.entries.clear();
      reverse_move_memory.entries.clear();
    }
    has_best = true;
    best_score = initial;
  }
  void phase_ended() {
    if (kind == AcceptorKind::GreatDeluge) has_water = false;  // great_deluge.rs:87-90
    if (kind == AcceptorKind::StepCountingHillClimbing) {      // step_counting.rs:92-95
      has_best = false;
      steps_since_improvement = 0;
    }
    if (kind == AcceptorKind::TabuSearch) {  // tabu_search.rs:196-202
      entity_memory.entries.clear();
      value_memory.entries.clear();
      move_memory.entries.clear();
      reverse_move_memory.entries.clear();
      has_best = false;
    }
  }
  bool is_tabu(const TabuSignature& sig) const {  // tabu_search.rs:150-162
    for (uint64_t e : sig.entity_ids)
      if (entity_memory.contains({sig.scope, e})) return true;
    for (uint64_t v : sig.value_ids)
      if (value_memory.contains({sig.scope, v})) return true;
    return move_memory.contains(sig.move_id) || reverse_move_memory.contains(sig.move_id);
  }
  bool is_accepted(Sc last, Sc mv, const TabuSignature* sig = nullptr) {
    switch (kind) {
      case AcceptorKind::AcceptAll: return true;
      case AcceptorKind::HillClimbing: return mv > last;  // hill_climbing.rs:33-42
      case AcceptorKind::LateAcceptance: return mv >= last || mv >= history[history_idx];
      case AcceptorKind::GreatDeluge:  // great_deluge.rs:53-68
        if (mv > last) return true;
        return has_water ? mv >= water_level : true;
      case AcceptorKind::StepCountingHillClimbing:  // step_counting.rs:56-67
        if (mv > last) return true;
        return steps_since_improvement < step_count_limit;
      case AcceptorKind::DiversifiedLateAcceptance: {  // diversified_late_acceptance.rs:72-101
        if (mv >= last) return true;
        if (mv >= history[history_idx]) return true;
        if (has_best) {
          Sc threshold = best_score - best_score.abs().multiply(tolerance);
          if (mv >= threshold) return true;
        }
        return false;
      }
      case AcceptorKind::TabuSearch: {  // tabu_search.rs:170-186
        if (!sig) throw std::logic_error("tabu search requires move signatures");
        bool aspirational = has_best && aspiration_enabled && mv > best_score;
        return aspirational || !is_tabu(*sig);
      }
      case AcceptorKind::SimulatedAnnealing: {
        if (mv >= last) return true;
        int lvl = 0;
        for (; lvl < Sc::levels; ++lvl)
          if (mv.level(lvl) != last.level(lvl)) break;
        double delta = (double)(mv.level(lvl) - last.level(lvl));  // < 0
        if (never_accept_hard_regression && Sc::level_is_hard(lvl)) return false;  // :353-357
        if (!calibrated) {  // CalibrationState::record / temperatures, :57-88,359-366
          sample_sum[lvl] += std::fabs(delta);
          sample_count[lvl] += 1;
          size_t total = 0;
          for (auto c : sample_count) total += c;
          if (total < calibration_samples) return false;
          for (int l = 0; l < Sc::levels; ++l) {
            double t = sample_count[l] ? (sample_sum[l] / (double)sample_count[l]) / -std::log(target_acceptance)
                                       : fallback_temperature;
            level_temperature[l] = std::max(t, fallback_temperature);
          }
          calibrated = true;
        }
        double T = level_temperature[lvl];
        if (T <= hill_climbing_temperature) return false;
        return uniform() < std::exp(delta / T);
      }
    }
    return false;
  }
  void step_ended(Sc step_score, const TabuSignature* accepted = nullptr) {
    if (kind == AcceptorKind::LateAcceptance || kind == AcceptorKind::DiversifiedLateAcceptance) {
      history[history_idx] = step_score;
      history_idx = (history_idx + 1) % history.size();
    }
    if (kind == AcceptorKind::SimulatedAnnealing && calibrated)
      for (auto& t : level_temperature) t = std::max(t * decay, hill_climbing_temperature);
    if (kind == AcceptorKind::GreatDeluge && has_water)  // great_deluge.rs:76-85
      water_level = water_level + initial_abs_score.multiply(rain_speed);
    if (kind == AcceptorKind::StepCountingHillClimbing) {  // step_counting.rs:74-90
      if (!has_best || step_score > best_score) {
        has_best = true;
        best_score = step_score;
        steps_since_improvement = 0;
      } else {
        steps_since_improvement += 1;
      }
      return;
    }
    if (kind == AcceptorKind::TabuSearch && accepted) {  // tabu_search.rs:214-225
      for (uint64_t e : accepted->entity_ids) entity_memory.record({accepted->scope, e});
      for (uint64_t v : accepted->value_ids) value_memory.record({accepted->scope, v});
      move_memory.record(accepted->move_id);
      reverse_move_memory.record(accepted->undo_move_id);
    }
    if (!has_best || step_score > best_score) {  // DLA :126-133, tabu :227-234
      has_best = true;
      best_score = step_score;
    }
  }
};

// Replay of phase/candidates.rs:66-282 given per-candidate (doable, score) in pull order.
template <class Sc>
struct StepOutcome {
  bool has_winner = false;
  size_t winner = 0;
  Sc winner_score = Sc::zero();
  uint64_t moves_evaluated = 0, score_calculations = 0, moves_accepted = 0;
};

struct NoSignatures {
  const TabuSignature* operator()(size_t) const { return nullptr; }
};

// tabu_signature of the scalar and list moves against the CURRENT working solution
// (change.rs:189-220, swap.rs:228-260, list_kernel/change.rs:155-204, list_kernel/swap.rs:110-155).
// Value ids: the reference hashes the Debug form of the value; raw index (None = NONE_ID) here.
template <class S, class Sc>
TabuSignature tabu_signature(const Move& m, ScoreDirector<S, Sc>& dir, uint64_t variable_id = 0) {
  S& s = dir.working;
  auto& ac = dir.access;
  auto vid = [](OptVal v) { return v ? (uint64_t)*v : TABU_NONE_ID; };
  TabuSignature sig;
  sig.scope = ((uint64_t)m.desc << 32) | variable_id;
  switch (m.kind) {
    case Move::Change: {
      const uint64_t from = vid(ac.get(s, m.desc, m.a)), to = vid(m.to);
      sig.move_id = {(uint64_t)m.desc, variable_id, (uint64_t)m.a, from, to};
      sig.undo_move_id = {(uint64_t)m.desc, variable_id, (uint64_t)m.a, to, from};
      sig.entity_ids = {(uint64_t)m.a};
      sig.value_ids = {to};
      return sig;
    }
    case Move::Swap: {
      const uint64_t lv = vid(ac.get(s, m.desc, m.a)), rv = vid(ac.get(s, m.desc, m.b));
      const uint64_t lo = std::min<uint64_t>(m.a, m.b), hi = std::max<uint64_t>(m.a, m.b);
      sig.move_id = {TABU_OP_SWAP, (uint64_t)m.desc, variable_id, lo, hi};
      sig.undo_move_id = sig.move_id;
      sig.entity_ids = {(uint64_t)m.a, (uint64_t)m.b};
      sig.value_ids = {rv, lv};
      return sig;
    }
    case Move::ListChange: {
      const auto& src = ac.list(s, m.desc, m.a);
      const uint64_t moved = m.b < src.size() ? (uint64_t)src[m.b] : TABU_NONE_ID;
      const uint64_t adj = adjusted_destination(m);
      sig.move_id = {(uint64_t)m.desc, variable_id, (uint64_t)m.a, (uint64_t)m.b, (uint64_t)m.c, adj, moved};
      sig.undo_move_id = {(uint64_t)m.desc, variable_id, (uint64_t)m.c, adj, (uint64_t)m.a, (uint64_t)m.b, moved};
      sig.entity_ids = {(uint64_t)m.a};
      if (m.a != m.c) sig.entity_ids.push_back((uint64_t)m.c);
      sig.value_ids = {moved};
      return sig;
    }
    case Move::ListSwap: {
      const auto& l1 = ac.list(s, m.desc, m.a);
      const uint64_t v1 = m.b < l1.size() ? (uint64_t)l1[m.b] : TABU_NONE_ID;
      const auto& l2 = ac.list(s, m.desc, m.c);
      const uint64_t v2 = m.d < l2.size() ? (uint64_t)l2[m.d] : TABU_NONE_ID;
      std::pair<uint64_t, uint64_t> p1{m.a, m.b}, p2{m.c, m.d};
      if (p2 < p1) std::swap(p1, p2);
      sig.move_id = {0xF000000000000003ull, (uint64_t)m.desc, variable_id, p1.first, p1.second, p2.first, p2.second};
      sig.undo_move_id = sig.move_id;
      sig.entity_ids = {(uint64_t)m.a};
      if (m.a != m.c) sig.entity_ids.push_back((uint64_t)m.c);
      sig.value_ids = {v2, v1};
      return sig;
    }
    case Move::ListReverse: {  // reverse.rs:60-98
      const auto& l = ac.list(s, m.desc, m.a);
      for (size_t p = m.b; p < m.c && p < l.size(); ++p) sig.value_ids.push_back((uint64_t)l[p]);
      sig.move_id = {0xF000000000000004ull, (uint64_t)m.desc, variable_id, (uint64_t)m.a, (uint64_t)m.b, (uint64_t)m.c};
      sig.undo_move_id = sig.move_id;
      sig.entity_ids = {(uint64_t)m.a};
      return sig;
    }
    case Move::SublistChange: {  // sublist_change.rs:127-196
      const auto& l = ac.list(s, m.desc, m.a);
      std::vector<uint64_t> moved;
      for (size_t p = m.b; p < m.c && p < l.size(); ++p) moved.push_back((uint64_t)l[p]);
      const Move inv = sublist_change_inverse(m);
      sig.move_id = {(uint64_t)m.desc, variable_id, (uint64_t)m.a, (uint64_t)m.b, (uint64_t)m.c, (uint64_t)m.d, (uint64_t)m.e};
      sig.undo_move_id = {(uint64_t)m.desc, variable_id, (uint64_t)inv.a, (uint64_t)inv.b, (uint64_t)inv.c, (uint64_t)inv.d, (uint64_t)inv.e};
      sig.move_id.insert(sig.move_id.end(), moved.begin(), moved.end());
      sig.undo_move_id.insert(sig.undo_move_id.end(), moved.begin(), moved.end());
      sig.entity_ids = {(uint64_t)m.a};
      if (m.a != m.d) sig.entity_ids.push_back((uint64_t)m.d);
      sig.value_ids = moved;
      return sig;
    }
    case Move::SublistSwap: {  // sublist_swap.rs:172-253
      const auto& l1 = ac.list(s, m.desc, m.a);
      const auto& l2 = ac.list(s, m.desc, m.d);
      std::vector<uint64_t> v1, v2;
      for (size_t p = m.b; p < m.c && p < l1.size(); ++p) v1.push_back((uint64_t)l1[p]);
      for (size_t p = m.e; p < m.f && p < l2.size(); ++p) v2.push_back((uint64_t)l2[p]);
      const Move inv = sublist_swap_inverse(m);
      sig.move_id = {(uint64_t)m.desc, variable_id, (uint64_t)m.a, (uint64_t)m.b, (uint64_t)m.c, (uint64_t)m.d, (uint64_t)m.e, (uint64_t)m.f};
      sig.undo_move_id = {(uint64_t)m.desc, variable_id, (uint64_t)inv.a, (uint64_t)inv.b, (uint64_t)inv.c, (uint64_t)inv.d, (uint64_t)inv.e, (uint64_t)inv.f};
      sig.move_id.insert(sig.move_id.end(), v1.begin(), v1.end());
      sig.move_id.insert(sig.move_id.end(), v2.begin(), v2.end());
      sig.undo_move_id.insert(sig.undo_move_id.end(), v2.begin(), v2.end());
      sig.undo_move_id.insert(sig.undo_move_id.end(), v1.begin(), v1.end());
      sig.entity_ids = {(uint64_t)m.a};
      if (m.a != m.d) sig.entity_ids.push_back((uint64_t)m.d);
      sig.value_ids = v1;
      sig.value_ids.insert(sig.value_ids.end(), v2.begin(), v2.end());
      return sig;
    }
    default: throw std::logic_error("tabu_signature: move kind not restated");
  }
}

template <class Sc, class GetEval, class GetSig = NoSignatures>
StepOutcome<Sc> replay_step(size_t n_candidates, GetEval eval, Sc best_score, Sc last_step_score,
                            uint64_t step_seed, Forager<Sc>& forager, Acceptor<Sc>& acceptor,
                            GetSig signature = GetSig()) {
  StepOutcome<Sc> out;
  forager.step_started(best_score, last_step_score, step_seed);
  for (size_t i = 0; i < n_candidates; ++i) {
    if (forager.is_quit_early()) break;
    CandidateEvaluation<Sc> ev = eval(i);
    out.moves_evaluated += 1;
    if (ev.kind == EvalKind::NotDoable) continue;
    out.score_calculations += 1;
    if (ev.kind != EvalKind::Scored) continue;
    if (acceptor.is_accepted(last_step_score, ev.score, signature(i))) {
      out.moves_accepted += 1;
      forager.add_move_index(i, ev.score);
    }
  }
  if (forager.best.has) {
    out.has_winner = true;
    out.winner = forager.best.index;
    out.winner_score = forager.best.score;
  }
  return out;
}

}  // namespace sfo
