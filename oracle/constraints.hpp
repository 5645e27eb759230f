// ORACLE — TEST INFRASTRUCTURE ONLY (see score.hpp header). Parity status: pinned against the
// reference's own known-answer tests re-encoded in oracle/kat_main.cpp (file:line cited there).
//
// Closure-based CPU restatement of solverforge-scoring's retained incremental constraints.
// This is BOTH the checker for the CUDA path and the "reference-faithful" CPU baseline: it
// keeps the same retained state as the reference (match rows + per-side buckets for joins,
// per-key B counts + A score totals for exists, group accumulators + cached per-group scores)
// and is driven by the same retract -> mutate -> insert protocol.
//
//   IncrementalConstraint / ConstraintSet   solverforge-scoring/src/api/constraint_set/incremental.rs:31-107,152-212,339-408
//   ChangeSource::assert_localizes          stream/collection_extract.rs:51-94
//   UniConstraint                           constraint/incremental.rs:97-156
//   CrossBiConstraint                       constraint/cross_bi_incremental/{state.rs:215-460, incremental.rs:27-137}
//   SelfJoinBiConstraint                    constraint/nary_incremental/bi.rs:78-206
//   SelfJoinNaryConstraint (tri/quad/penta) constraint/nary_incremental/higher_arity/shared.rs:114-380
//   ExistsConstraint                        constraint/exists.rs:126-417 (+ exists/key_state.rs:95-253)
//   GroupedConstraint                       constraint/grouped/{state.rs:102-261, scorer.rs:46-152}
//   collectors count / sum / load_balance   stream/collector/{count.rs, sum.rs, load_balance.rs:104-240}
//   CrossComplementedGroupedConstraint      constraint/cross_complemented_grouped/{state.rs:173-440, updates.rs:32-249, view.rs:24-203}
//                                           + constraint/grouped/complemented_scorer.rs:63-120
//   ComplementedGroupedConstraint (uni)     constraint/complemented/{helpers.rs:29-80, incremental.rs:29-49}
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <set>
#include <array>
#include <utility>
#include <vector>

#include "score.hpp"

namespace sfo {

enum class Impact { Penalty, Reward };

struct ChangeSource {
  enum Kind { Unknown, Static, Descriptor } kind = Unknown;
  size_t index = 0;
  static ChangeSource Stat() { return {Static, 0}; }
  static ChangeSource Desc(size_t i) { return {Descriptor, i}; }
  bool owns(size_t d) const { return kind == Descriptor && index == d; }
  bool same_index_domain(ChangeSource o) const {
    return kind == Descriptor && o.kind == Descriptor && index == o.index;
  }
  // collection_extract.rs:83-93: Descriptor(i)==d -> react; Static -> ignore; Unknown -> panic.
  bool assert_localizes(size_t d, const std::string& name) const {
    if (owns(d)) return true;
    if (kind == Unknown) throw std::logic_error("constraint `" + name + "` cannot localize entity indexes");
    return false;
  }
};

struct PairHash {
  size_t operator()(const std::pair<size_t, size_t>& p) const {
    uint64_t x = (uint64_t)p.first * 0x9E3779B97F4A7C15ull ^ ((uint64_t)p.second + 0x7F4A7C15ull);
    x ^= x >> 32;
    return (size_t)(x * 0xD6E8FEB86659FD93ull);
  }
};

template <class S, class Sc>
struct IncrementalConstraint {
  std::string name;
  bool is_hard = false;
  virtual ~IncrementalConstraint() = default;
  virtual Sc evaluate(const S& s) const = 0;
  virtual size_t match_count(const S& s) const = 0;
  virtual Sc initialize(const S& s) = 0;
  virtual Sc on_insert(const S& s, size_t entity_index, size_t descriptor_index) = 0;
  virtual Sc on_retract(const S& s, size_t entity_index, size_t descriptor_index) = 0;
  virtual void reset() = 0;
};

// api/constraint_set/incremental.rs:339-408 — the tuple impls sum every member.
template <class S, class Sc>
struct ConstraintSet {
  std::vector<std::unique_ptr<IncrementalConstraint<S, Sc>>> items;
  template <class C>
  C* add(std::unique_ptr<C> c) {
    C* raw = c.get();
    items.emplace_back(std::move(c));
    return raw;
  }
  Sc evaluate_all(const S& s) const {
    Sc t = Sc::zero();
    for (auto& c : items) t = t + c->evaluate(s);
    return t;
  }
  Sc initialize_all(const S& s) {
    Sc t = Sc::zero();
    for (auto& c : items) t = t + c->initialize(s);
    return t;
  }
  Sc on_insert_all(const S& s, size_t e, size_t d) {
    Sc t = Sc::zero();
    for (auto& c : items) t = t + c->on_insert(s, e, d);
    return t;
  }
  Sc on_retract_all(const S& s, size_t e, size_t d) {
    Sc t = Sc::zero();
    for (auto& c : items) t = t + c->on_retract(s, e, d);
    return t;
  }
  void reset_all() {
    for (auto& c : items) c->reset();
  }
  size_t constraint_count() const { return items.size(); }
};

template <class Sc>
inline Sc signed_weight(Impact impact, Sc base) {
  return impact == Impact::Penalty ? -base : base;
}

// A source = extractor + source-level predicate (`contains`) + ChangeSource.
template <class S, class A>
struct Source {
  const std::vector<A>& (*extract)(const S&);
  ChangeSource change = ChangeSource{};
  bool (*contains)(const S&, const A&) = nullptr;  // nullptr => true
  bool has(const S& s, const A& a) const { return contains == nullptr || contains(s, a); }
};

// ---------------------------------------------------------------------------------------------
// Uni: sum_e [filter(s,e)] * +-w(e)                          constraint/incremental.rs:97-156
template <class S, class A, class Sc, class F, class W>
struct UniConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> src;
  Impact impact;
  F filter;  // (const S&, const A&) -> bool
  W weight;  // (const A&) -> Sc
  UniConstraint(std::string n, Impact i, Source<S, A> s, F f, W w, bool hard)
      : src(s), impact(i), filter(std::move(f)), weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  Sc evaluate(const S& s) const override {
    Sc t = Sc::zero();
    for (auto& e : src.extract(s))
      if (filter(s, e)) t = t + signed_weight(impact, weight(e));
    return t;
  }
  size_t match_count(const S& s) const override {
    size_t n = 0;
    for (auto& e : src.extract(s)) n += filter(s, e) ? 1 : 0;
    return n;
  }
  Sc initialize(const S& s) override { return evaluate(s); }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    auto& es = src.extract(s);
    if (idx >= es.size()) return Sc::zero();
    return filter(s, es[idx]) ? signed_weight(impact, weight(es[idx])) : Sc::zero();
  }
  Sc on_retract(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    auto& es = src.extract(s);
    if (idx >= es.size()) return Sc::zero();
    return filter(s, es[idx]) ? -signed_weight(impact, weight(es[idx])) : Sc::zero();
  }
  void reset() override {}
};

// ---------------------------------------------------------------------------------------------
// Shared retained join state: matches map + match rows + per-side buckets + per-key indexes.
// Payload P is the per-row data (a score for CrossBi, (group, retraction) for grouped joins).
template <class K, class KH, class P>
struct JoinRows {
  struct Row {
    std::pair<size_t, size_t> pair;
    P payload;
    size_t a_pos, b_pos;
  };
  std::unordered_map<std::pair<size_t, size_t>, size_t, PairHash> matches;
  std::vector<Row> rows;
  std::unordered_map<size_t, std::vector<size_t>> a_to, b_to;
  std::unordered_map<K, std::vector<size_t>, KH> a_by_key, b_by_key;
  std::unordered_map<size_t, K> a_index_to_key, b_index_to_key;

  void clear() {
    matches.clear();
    rows.clear();
    a_to.clear();
    b_to.clear();
    a_by_key.clear();
    b_by_key.clear();
    a_index_to_key.clear();
    b_index_to_key.clear();
  }
  void push_row(size_t a, size_t b, P payload) {
    size_t row_idx = rows.size();
    auto& ab = a_to[a];
    size_t a_pos = ab.size();
    ab.push_back(row_idx);
    auto& bb = b_to[b];
    size_t b_pos = bb.size();
    bb.push_back(row_idx);
    rows.push_back(Row{{a, b}, std::move(payload), a_pos, b_pos});
    matches.emplace(std::make_pair(a, b), row_idx);
  }
  static void bucket_remove(std::unordered_map<size_t, std::vector<size_t>>& side, std::vector<Row>& rows,
                            size_t idx, size_t row_idx, size_t pos, bool a_side) {
    auto it = side.find(idx);
    if (it == side.end()) return;
    auto& v = it->second;
    size_t rp = pos;
    if (!(rp < v.size() && v[rp] == row_idx)) {
      auto f = std::find(v.begin(), v.end(), row_idx);
      if (f == v.end()) return;
      rp = (size_t)(f - v.begin());
    }
    v[rp] = v.back();
    v.pop_back();
    if (rp < v.size()) {
      if (a_side) rows[v[rp]].a_pos = rp; else rows[v[rp]].b_pos = rp;
    }
    if (v.empty()) side.erase(it);
  }
  // state.rs:299-322 (swap_remove + fix-up of the moved row)
  P remove_row(size_t row_idx) {
    Row row = rows[row_idx];
    matches.erase(row.pair);
    bucket_remove(a_to, rows, row.pair.first, row_idx, row.a_pos, true);
    bucket_remove(b_to, rows, row.pair.second, row_idx, row.b_pos, false);
    size_t last = rows.size() - 1;
    if (row_idx != last) rows[row_idx] = rows[last];
    rows.pop_back();
    if (row_idx != last) {
      Row& moved = rows[row_idx];
      matches[moved.pair] = row_idx;
      auto ia = a_to.find(moved.pair.first);
      if (ia != a_to.end()) ia->second[moved.a_pos] = row_idx;
      auto ib = b_to.find(moved.pair.second);
      if (ib != b_to.end()) ib->second[moved.b_pos] = row_idx;
    }
    return row.payload;
  }
  static void key_bucket_remove(std::unordered_map<K, std::vector<size_t>, KH>& m, const K& key, size_t idx) {
    auto it = m.find(key);
    if (it == m.end()) return;
    auto& v = it->second;
    auto f = std::find(v.begin(), v.end(), idx);
    if (f != v.end()) {
      *f = v.back();
      v.pop_back();
    }
    if (v.empty()) m.erase(it);
  }
};

// ---------------------------------------------------------------------------------------------
// Cross-bi keyed join. Predicate joins use a constant key (stream/join_target.rs:83-110).
template <class S, class A, class B, class K, class Sc, class KA, class KB, class F, class W,
          class KH = std::hash<K>>
struct CrossBiConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> sa;
  Source<S, B> sb;
  Impact impact;
  KA key_a;
  KB key_b;
  F filter;  // (const S&, const A&, const B&, size_t ia, size_t ib) -> bool
  W weight;  // (const S&, const A&, const B&, size_t ia, size_t ib) -> Sc
  JoinRows<K, KH, Sc> st;

  CrossBiConstraint(std::string n, Impact i, Source<S, A> a, Source<S, B> b, KA ka, KB kb, F f, W w, bool hard)
      : sa(a), sb(b), impact(i), key_a(std::move(ka)), key_b(std::move(kb)), filter(std::move(f)),
        weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  Sc score_of(const S& s, const std::vector<A>& ea, const std::vector<B>& eb, size_t ia, size_t ib) const {
    return signed_weight(impact, weight(s, ea[ia], eb[ib], ia, ib));
  }
  std::unordered_map<K, std::vector<size_t>, KH> b_index_for(const S& s, const std::vector<B>& eb) const {
    std::unordered_map<K, std::vector<size_t>, KH> m;
    for (size_t i = 0; i < eb.size(); ++i)
      if (sb.has(s, eb[i])) m[key_b(eb[i])].push_back(i);
    return m;
  }
  Sc evaluate(const S& s) const override {
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    auto idx = b_index_for(s, eb);
    Sc t = Sc::zero();
    for (size_t ia = 0; ia < ea.size(); ++ia) {
      if (!sa.has(s, ea[ia])) continue;
      auto it = idx.find(key_a(ea[ia]));
      if (it == idx.end()) continue;
      for (size_t ib : it->second)
        if (filter(s, ea[ia], eb[ib], ia, ib)) t = t + score_of(s, ea, eb, ia, ib);
    }
    return t;
  }
  size_t match_count(const S& s) const override {
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    auto idx = b_index_for(s, eb);
    size_t n = 0;
    for (size_t ia = 0; ia < ea.size(); ++ia) {
      if (!sa.has(s, ea[ia])) continue;
      auto it = idx.find(key_a(ea[ia]));
      if (it == idx.end()) continue;
      for (size_t ib : it->second) n += filter(s, ea[ia], eb[ib], ia, ib) ? 1 : 0;
    }
    return n;
  }
  Sc add_match(const S& s, const std::vector<A>& ea, const std::vector<B>& eb, size_t ia, size_t ib) {
    if (st.matches.count({ia, ib})) return Sc::zero();
    if (!sa.has(s, ea[ia]) || !sb.has(s, eb[ib])) return Sc::zero();
    if (!filter(s, ea[ia], eb[ib], ia, ib)) return Sc::zero();
    Sc sc = score_of(s, ea, eb, ia, ib);
    st.push_row(ia, ib, sc);
    return sc;
  }
  Sc initialize(const S& s) override {
    reset();
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    for (size_t i = 0; i < ea.size(); ++i) {
      if (!sa.has(s, ea[i])) continue;
      K k = key_a(ea[i]);
      st.a_index_to_key.emplace(i, k);
      st.a_by_key[k].push_back(i);
    }
    for (size_t i = 0; i < eb.size(); ++i) {
      if (!sb.has(s, eb[i])) continue;
      K k = key_b(eb[i]);
      st.b_index_to_key.emplace(i, k);
      st.b_by_key[k].push_back(i);
    }
    Sc t = Sc::zero();
    for (size_t ia = 0; ia < ea.size(); ++ia) {
      if (!sa.has(s, ea[ia])) continue;
      auto it = st.b_by_key.find(key_a(ea[ia]));
      if (it == st.b_by_key.end()) continue;
      std::vector<size_t> bs = it->second;  // the reference clones the bucket (state.rs:392)
      for (size_t ib : bs) t = t + add_match(s, ea, eb, ia, ib);
    }
    return t;
  }
  Sc insert_a(const S& s, const std::vector<A>& ea, const std::vector<B>& eb, size_t ia) {
    if (ia >= ea.size()) return Sc::zero();
    if (!sa.has(s, ea[ia])) return Sc::zero();
    K k = key_a(ea[ia]);
    st.a_index_to_key[ia] = k;
    st.a_by_key[k].push_back(ia);
    auto it = st.b_by_key.find(k);
    std::vector<size_t> bs = it == st.b_by_key.end() ? std::vector<size_t>{} : it->second;
    Sc t = Sc::zero();
    for (size_t ib : bs) t = t + add_match(s, ea, eb, ia, ib);
    return t;
  }
  Sc insert_b(const S& s, const std::vector<A>& ea, const std::vector<B>& eb, size_t ib) {
    if (ib >= eb.size()) return Sc::zero();
    if (!sb.has(s, eb[ib])) return Sc::zero();
    K k = key_b(eb[ib]);
    st.b_index_to_key[ib] = k;
    st.b_by_key[k].push_back(ib);
    auto it = st.a_by_key.find(k);
    std::vector<size_t> as = it == st.a_by_key.end() ? std::vector<size_t>{} : it->second;
    Sc t = Sc::zero();
    for (size_t ia : as) t = t + add_match(s, ea, eb, ia, ib);
    return t;
  }
  Sc retract_a(size_t ia) {
    auto ik = st.a_index_to_key.find(ia);
    if (ik != st.a_index_to_key.end()) {
      st.key_bucket_remove(st.a_by_key, ik->second, ia);
      st.a_index_to_key.erase(ik);
    }
    Sc t = Sc::zero();
    for (;;) {
      auto it = st.a_to.find(ia);
      if (it == st.a_to.end() || it->second.empty()) break;
      t = t + (-st.remove_row(it->second.back()));
    }
    return t;
  }
  Sc retract_b(size_t ib) {
    auto ik = st.b_index_to_key.find(ib);
    if (ik != st.b_index_to_key.end()) {
      st.key_bucket_remove(st.b_by_key, ik->second, ib);
      st.b_index_to_key.erase(ik);
    }
    Sc t = Sc::zero();
    for (;;) {
      auto it = st.b_to.find(ib);
      if (it == st.b_to.end() || it->second.empty()) break;
      t = t + (-st.remove_row(it->second.back()));
    }
    return t;
  }
  // incremental.rs:89-137: A side first, then B side, for both insert and retract.
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    bool ac = sa.change.assert_localizes(d, this->name);
    bool bc = sb.change.assert_localizes(d, this->name);
    Sc t = Sc::zero();
    if (!ac && !bc) return t;
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    if (ac) t = t + insert_a(s, ea, eb, idx);
    if (bc) t = t + insert_b(s, ea, eb, idx);
    return t;
  }
  Sc on_retract(const S&, size_t idx, size_t d) override {
    bool ac = sa.change.assert_localizes(d, this->name);
    bool bc = sb.change.assert_localizes(d, this->name);
    Sc t = Sc::zero();
    if (!ac && !bc) return t;
    if (ac) t = t + retract_a(idx);
    if (bc) t = t + retract_b(idx);
    return t;
  }
  void reset() override { st.clear(); }
};

// ---------------------------------------------------------------------------------------------
// Keyed self-join over one collection: unordered pairs low < high inside a key bucket.
//                                                         constraint/nary_incremental/bi.rs:78-206
template <class S, class A, class K, class Sc, class KF, class F, class W, class KH = std::hash<K>>
struct SelfJoinBiConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> src;
  Impact impact;
  KF key_fn;
  F filter;  // (const S&, const A& low, const A& high, size_t low, size_t high) -> bool
  W weight;  // (const S&, const A& low, const A& high) -> Sc
  std::unordered_map<size_t, std::unordered_set<size_t>> entity_to_matches;  // entity -> partner set
  std::unordered_set<std::pair<size_t, size_t>, PairHash> matches;
  std::unordered_map<K, std::vector<size_t>, KH> key_to_indices;
  std::unordered_map<size_t, K> index_to_key;

  SelfJoinBiConstraint(std::string n, Impact i, Source<S, A> s, KF kf, F f, W w, bool hard)
      : src(s), impact(i), key_fn(std::move(kf)), filter(std::move(f)), weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  Sc evaluate(const S& s) const override {
    auto& es = src.extract(s);
    std::unordered_map<K, std::vector<size_t>, KH> idx;
    for (size_t i = 0; i < es.size(); ++i)
      if (src.has(s, es[i])) idx[key_fn(es[i])].push_back(i);
    Sc t = Sc::zero();
    for (auto& kv : idx) {
      auto& v = kv.second;
      for (size_t x = 0; x < v.size(); ++x)
        for (size_t y = x + 1; y < v.size(); ++y) {
          size_t lo = std::min(v[x], v[y]), hi = std::max(v[x], v[y]);
          if (filter(s, es[lo], es[hi], lo, hi)) t = t + signed_weight(impact, weight(s, es[lo], es[hi]));
        }
    }
    return t;
  }
  size_t match_count(const S& s) const override {
    auto& es = src.extract(s);
    size_t n = 0;
    for (size_t lo = 0; lo < es.size(); ++lo)
      for (size_t hi = lo + 1; hi < es.size(); ++hi)
        if (src.has(s, es[lo]) && src.has(s, es[hi]) && key_fn(es[lo]) == key_fn(es[hi]) &&
            filter(s, es[lo], es[hi], lo, hi))
          ++n;
    return n;
  }
  Sc insert_entity(const S& s, const std::vector<A>& es, size_t idx) {
    if (idx >= es.size() || !src.has(s, es[idx])) return Sc::zero();
    K k = key_fn(es[idx]);
    index_to_key[idx] = k;
    auto& bucket = key_to_indices[k];
    Sc t = Sc::zero();
    for (size_t other : bucket) {
      if (other == idx) continue;
      size_t lo = std::min(idx, other), hi = std::max(idx, other);
      if (matches.count({lo, hi})) continue;
      if (!filter(s, es[lo], es[hi], lo, hi)) continue;
      matches.insert({lo, hi});
      entity_to_matches[lo].insert(hi);
      entity_to_matches[hi].insert(lo);
      t = t + signed_weight(impact, weight(s, es[lo], es[hi]));
    }
    bucket.push_back(idx);
    return t;
  }
  // bi.rs:158-162 — retract recomputes the weight from the CURRENT entities.
  Sc retract_entity(const S& s, const std::vector<A>& es, size_t idx) {
    auto ik = index_to_key.find(idx);
    if (ik != index_to_key.end()) {
      auto ib = key_to_indices.find(ik->second);
      if (ib != key_to_indices.end()) {
        auto& v = ib->second;
        auto f = std::find(v.begin(), v.end(), idx);
        if (f != v.end()) {
          *f = v.back();
          v.pop_back();
        }
        if (v.empty()) key_to_indices.erase(ib);
      }
      index_to_key.erase(ik);
    }
    Sc t = Sc::zero();
    auto it = entity_to_matches.find(idx);
    if (it == entity_to_matches.end()) return t;
    std::unordered_set<size_t> partners = std::move(it->second);
    entity_to_matches.erase(it);
    for (size_t other : partners) {
      size_t lo = std::min(idx, other), hi = std::max(idx, other);
      matches.erase({lo, hi});
      auto io = entity_to_matches.find(other);
      if (io != entity_to_matches.end()) {
        io->second.erase(idx);
        if (io->second.empty()) entity_to_matches.erase(io);
      }
      if (lo < es.size() && hi < es.size()) t = t + (-signed_weight(impact, weight(s, es[lo], es[hi])));
    }
    return t;
  }
  Sc initialize(const S& s) override {
    reset();
    auto& es = src.extract(s);
    Sc t = Sc::zero();
    for (size_t i = 0; i < es.size(); ++i) t = t + insert_entity(s, es, i);
    return t;
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    return insert_entity(s, src.extract(s), idx);
  }
  Sc on_retract(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    return retract_entity(s, src.extract(s), idx);
  }
  void reset() override {
    entity_to_matches.clear();
    matches.clear();
    key_to_indices.clear();
    index_to_key.clear();
  }
};

// ---------------------------------------------------------------------------------------------
// Keyed self-join of arity N = 3 (tri), 4 (quad), 5 (penta): index-ordered tuples a < b < c .. inside a key
// bucket.                         constraint/nary_incremental/higher_arity/{shared.rs:114-380, tri.rs, quad.rs, penta.rs}
// A tuple is its ascending index list padded with SIZE_MAX; `filter(s, es, tuple)` and `weight(s, es, tuple)`.
using NaryTuple = std::array<size_t, 5>;
struct NaryTupleHash {
  size_t operator()(const NaryTuple& t) const {
    size_t h = 0xcbf29ce484222325ull;
    for (size_t v : t) h = (h ^ v) * 0x100000001b3ull;
    return h;
  }
};
template <class S, class A, class K, class Sc, class KF, class F, class W, class KH = std::hash<K>>
struct SelfJoinNaryConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> src;
  Impact impact;
  size_t arity;
  KF key_fn;
  F filter;
  W weight;
  std::unordered_set<NaryTuple, NaryTupleHash> matches;
  std::unordered_map<size_t, std::unordered_set<NaryTuple, NaryTupleHash>> entity_to_matches;
  std::unordered_map<K, std::set<size_t>, KH> key_to_indices;
  std::unordered_map<size_t, K> index_to_key;

  SelfJoinNaryConstraint(std::string n, Impact i, size_t ar, Source<S, A> s, KF kf, F f, W w, bool hard)
      : src(s), impact(i), arity(ar), key_fn(std::move(kf)), filter(std::move(f)), weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  // every ascending combination of `need` members of `pool` starting at position `from`, appended to `cur`
  template <class V>
  static void combos(const std::vector<size_t>& pool, size_t from, size_t need, std::vector<size_t>& cur, V&& visit) {
    if (need == 0) {
      visit(cur);
      return;
    }
    for (size_t p = from; p + need <= pool.size(); ++p) {
      cur.push_back(pool[p]);
      combos(pool, p + 1, need - 1, cur, visit);
      cur.pop_back();
    }
  }
  static NaryTuple make_tuple(std::vector<size_t> v) {
    std::sort(v.begin(), v.end());
    NaryTuple t;
    t.fill(SIZE_MAX);
    for (size_t i = 0; i < v.size(); ++i) t[i] = v[i];
    return t;
  }
  Sc evaluate(const S& s) const override {  // shared.rs:281-304
    auto& es = src.extract(s);
    std::unordered_map<K, std::vector<size_t>, KH> idx;
    for (size_t i = 0; i < es.size(); ++i) idx[key_fn(es[i])].push_back(i);
    Sc t = Sc::zero();
    std::vector<size_t> cur;
    for (auto& kv : idx)
      combos(kv.second, 0, arity, cur, [&](const std::vector<size_t>& c) {
        NaryTuple tp = make_tuple(c);
        if (filter(s, es, tp)) t = t + signed_weight(impact, weight(s, es, tp));
      });
    return t;
  }
  size_t match_count(const S& s) const override { return matches.size(); }
  Sc insert_entity(const S& s, const std::vector<A>& es, size_t idx) {  // shared.rs:170-223
    if (idx >= es.size()) return Sc::zero();
    K k = key_fn(es[idx]);
    index_to_key[idx] = k;
    auto& bucket = key_to_indices[k];
    bucket.insert(idx);
    std::vector<size_t> others;
    for (size_t o : bucket)
      if (o != idx) others.push_back(o);
    Sc t = Sc::zero();
    std::vector<size_t> cur{idx};
    combos(others, 0, arity - 1, cur, [&](const std::vector<size_t>& c) {
      NaryTuple tp = make_tuple(c);
      if (matches.count(tp)) return;
      if (!filter(s, es, tp)) return;
      matches.insert(tp);
      for (size_t i = 0; i < arity; ++i) entity_to_matches[tp[i]].insert(tp);
      t = t + signed_weight(impact, weight(s, es, tp));
    });
    return t;
  }
  Sc retract_entity(const S& s, const std::vector<A>& es, size_t idx) {  // shared.rs:225-270
    auto ik = index_to_key.find(idx);
    if (ik != index_to_key.end()) {
      auto ib = key_to_indices.find(ik->second);
      if (ib != key_to_indices.end()) {
        ib->second.erase(idx);
        if (ib->second.empty()) key_to_indices.erase(ib);
      }
      index_to_key.erase(ik);
    }
    auto it = entity_to_matches.find(idx);
    if (it == entity_to_matches.end()) return Sc::zero();
    auto tuples = std::move(it->second);
    entity_to_matches.erase(it);
    Sc t = Sc::zero();
    for (const NaryTuple& tp : tuples) {
      matches.erase(tp);
      bool in_range = true;
      for (size_t i = 0; i < arity; ++i) {
        if (tp[i] != idx) {
          auto io = entity_to_matches.find(tp[i]);
          if (io != entity_to_matches.end()) {
            io->second.erase(tp);
            if (io->second.empty()) entity_to_matches.erase(io);
          }
        }
        in_range = in_range && tp[i] < es.size();
      }
      if (in_range) t = t + (-signed_weight(impact, weight(s, es, tp)));
    }
    return t;
  }
  Sc initialize(const S& s) override {
    reset();
    auto& es = src.extract(s);
    Sc t = Sc::zero();
    for (size_t i = 0; i < es.size(); ++i) t = t + insert_entity(s, es, i);
    return t;
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    return insert_entity(s, src.extract(s), idx);
  }
  Sc on_retract(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    return retract_entity(s, src.extract(s), idx);
  }
  void reset() override {
    matches.clear();
    entity_to_matches.clear();
    key_to_indices.clear();
    index_to_key.clear();
  }
};

// ---------------------------------------------------------------------------------------------
// Exists / NotExists with an optionally flattened B side.             constraint/exists.rs
enum class ExistenceMode { Exists, NotExists };

template <class S, class A, class P, class B, class K, class Sc, class KA, class KB, class FA, class FP,
          class Flatten, class W, class KH = std::hash<K>>
struct ExistsConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> sa;
  Source<S, P> sp;
  Impact impact;
  ExistenceMode mode;
  KA key_a;
  KB key_b;
  FA filter_a;       // (const S&, const A&) -> bool
  FP filter_parent;  // (const S&, const P&) -> bool
  Flatten flatten;   // (const P&) -> const std::vector<B>&   (SelfFlatten: vector of one)
  W weight;          // (const A&) -> Sc
  struct ASlot {
    std::optional<K> key;
    size_t bucket_pos = 0;
    Sc score = Sc::zero();
  };
  std::vector<ASlot> a_slots;
  // exists/key_state.rs — hashed storage variant (dense-Vec storage is an optimisation with
  // identical results: constraint/tests/exists_storage.rs).
  std::unordered_map<K, size_t, KH> b_counts;
  std::unordered_map<K, Sc, KH> a_score_totals;
  std::unordered_map<K, std::vector<size_t>, KH> a_buckets;

  ExistsConstraint(std::string n, Impact i, ExistenceMode m, Source<S, A> a, Source<S, P> p, KA ka, KB kb,
                   FA fa, FP fp, Flatten fl, W w, bool hard)
      : sa(a), sp(p), impact(i), mode(m), key_a(std::move(ka)), key_b(std::move(kb)),
        filter_a(std::move(fa)), filter_parent(std::move(fp)), flatten(std::move(fl)), weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  bool matches_count(size_t c) const { return mode == ExistenceMode::Exists ? c > 0 : c == 0; }
  size_t b_count(const K& k) const {
    auto it = b_counts.find(k);
    return it == b_counts.end() ? 0 : it->second;
  }
  Sc a_total(const K& k) const {
    auto it = a_score_totals.find(k);
    return it == a_score_totals.end() ? Sc::zero() : it->second;
  }
  std::unordered_map<K, size_t, KH> build_b_counts(const S& s) const {
    std::unordered_map<K, size_t, KH> m;
    for (auto& p : sp.extract(s)) {
      if (!sp.has(s, p) || !filter_parent(s, p)) continue;
      for (auto& item : flatten(p)) m[key_b(item)] += 1;
    }
    return m;
  }
  Sc evaluate(const S& s) const override {
    auto counts = build_b_counts(s);
    Sc t = Sc::zero();
    for (auto& a : sa.extract(s)) {
      if (!sa.has(s, a) || !filter_a(s, a)) continue;
      auto it = counts.find(key_a(a));
      if (matches_count(it == counts.end() ? 0 : it->second)) t = t + signed_weight(impact, weight(a));
    }
    return t;
  }
  size_t match_count(const S& s) const override {
    auto counts = build_b_counts(s);
    size_t n = 0;
    for (auto& a : sa.extract(s)) {
      if (!sa.has(s, a) || !filter_a(s, a)) continue;
      auto it = counts.find(key_a(a));
      n += matches_count(it == counts.end() ? 0 : it->second) ? 1 : 0;
    }
    return n;
  }
  Sc insert_a(const S& s, size_t idx) {
    auto& ea = sa.extract(s);
    if (idx >= ea.size()) return Sc::zero();
    if (a_slots.size() < ea.size()) a_slots.resize(ea.size());
    const A& a = ea[idx];
    if (!sa.has(s, a) || !filter_a(s, a)) {
      a_slots[idx] = ASlot{};
      return Sc::zero();
    }
    K k = key_a(a);
    auto& bucket = a_buckets[k];
    size_t pos = bucket.size();
    bucket.push_back(idx);
    Sc sc = signed_weight(impact, weight(a));
    a_score_totals[k] = a_total(k) + sc;
    Sc contribution = matches_count(b_count(k)) ? sc : Sc::zero();
    a_slots[idx] = ASlot{k, pos, sc};
    return contribution;
  }
  Sc retract_a(size_t idx) {
    if (idx >= a_slots.size()) return Sc::zero();
    ASlot slot = a_slots[idx];
    if (!slot.key) return Sc::zero();
    const K& k = *slot.key;
    Sc contribution = matches_count(b_count(k)) ? slot.score : Sc::zero();
    auto ib = a_buckets.find(k);
    if (ib != a_buckets.end()) {
      auto& v = ib->second;
      if (slot.bucket_pos < v.size() && v[slot.bucket_pos] == idx) {
        v[slot.bucket_pos] = v.back();
        v.pop_back();
        if (slot.bucket_pos < v.size()) a_slots[v[slot.bucket_pos]].bucket_pos = slot.bucket_pos;
      }
      if (v.empty()) a_buckets.erase(ib);
    }
    a_score_totals[k] = a_total(k) - slot.score;
    a_slots[idx] = ASlot{};
    return -contribution;
  }
  // exists.rs:259-270 — linear-dedupe list of (key, multiplicity) of one parent.
  std::vector<std::pair<K, size_t>> parent_key_counts(const S& s, size_t idx) const {
    std::vector<std::pair<K, size_t>> kc;
    auto& ps = sp.extract(s);
    if (idx >= ps.size()) return kc;
    if (!sp.has(s, ps[idx]) || !filter_parent(s, ps[idx])) return kc;
    for (auto& item : flatten(ps[idx])) {
      K k = key_b(item);
      bool found = false;
      for (auto& e : kc)
        if (e.first == k) {
          e.second += 1;
          found = true;
          break;
        }
      if (!found) kc.emplace_back(k, 1);
    }
    return kc;
  }
  Sc update_key_counts(const std::vector<std::pair<K, size_t>>& kc, bool insert) {
    Sc t = Sc::zero();
    for (auto& e : kc) {
      size_t old_c = b_count(e.first);
      size_t new_c = insert ? old_c + e.second : (old_c >= e.second ? old_c - e.second : 0);
      if (new_c == 0) b_counts.erase(e.first); else b_counts[e.first] = new_c;
      bool om = matches_count(old_c), nm = matches_count(new_c);
      if (om != nm) t = nm ? t + a_total(e.first) : t - a_total(e.first);
    }
    return t;
  }
  Sc initialize(const S& s) override {
    reset();
    b_counts = build_b_counts(s);
    size_t len = sa.extract(s).size();
    a_slots.assign(len, ASlot{});
    Sc t = Sc::zero();
    for (size_t i = 0; i < len; ++i) t = t + insert_a(s, i);
    return t;
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    bool ac = sa.change.assert_localizes(d, this->name);
    bool pc = sp.change.assert_localizes(d, this->name);
    Sc t = Sc::zero();
    if (pc) t = t + update_key_counts(parent_key_counts(s, idx), true);
    if (ac) t = t + insert_a(s, idx);
    return t;
  }
  Sc on_retract(const S& s, size_t idx, size_t d) override {
    bool ac = sa.change.assert_localizes(d, this->name);
    bool pc = sp.change.assert_localizes(d, this->name);
    bool same = sa.change.same_index_domain(sp.change) && ac && pc;
    Sc t = Sc::zero();
    if (same) {
      auto keys = parent_key_counts(s, idx);
      t = t + retract_a(idx);
      t = t + update_key_counts(keys, false);
      return t;
    }
    if (ac) t = t + retract_a(idx);
    if (pc) t = t + update_key_counts(parent_key_counts(s, idx), false);
    return t;
  }
  void reset() override {
    a_slots.clear();
    b_counts.clear();
    a_score_totals.clear();
    a_buckets.clear();
  }
};

// ---------------------------------------------------------------------------------------------
// Collectors (stream/collector/*): each has Value, Result, Accumulator{accumulate,retract,result,reset}.
struct CountAcc {  // count.rs: Result = usize
  using Value = char;
  using Result = size_t;
  using Retraction = char;
  size_t n = 0;
  Retraction accumulate(Value) { ++n; return 0; }
  void retract(Retraction) { n = n > 0 ? n - 1 : 0; }
  Result result() const { return n; }
  void reset() { n = 0; }
};
struct SumAcc {  // sum.rs: Result = T (i64 here)
  using Value = int64_t;
  using Result = int64_t;
  using Retraction = int64_t;
  int64_t sum = 0;
  Retraction accumulate(Value v) { sum = wadd(sum, v); return v; }
  void retract(Retraction v) { sum = wsub(sum, v); }
  Result result() const { return sum; }
  void reset() { sum = 0; }
};
// runs.rs:14-229 — consecutive_runs(index): the unique integer points of a group as maximal runs of
// consecutive values; duplicates raise a run's item_count, not its point_count.
struct Run {
  int64_t start = 0, end = 0;
  size_t point_count = 0, item_count = 0;
};
struct Runs {
  std::vector<Run> runs;
  size_t point_count = 0, item_count = 0;
};
struct RunsAcc {
  using Value = int64_t;
  using Result = Runs;
  using Retraction = int64_t;
  std::map<int64_t, size_t> points;  // BTreeMap: ordered
  size_t items = 0;
  Retraction accumulate(Value v) {
    points[v] += 1;
    items += 1;
    return v;
  }
  void retract(Retraction v) {  // :151-161
    auto it = points.find(v);
    if (it == points.end()) return;
    it->second = it->second > 0 ? it->second - 1 : 0;
    items = items > 0 ? items - 1 : 0;
    if (it->second == 0) points.erase(it);
  }
  Result result() const {  // runs_from_counts_and_item_count :179-229
    Runs out;
    out.point_count = points.size();
    out.item_count = items;
    bool open = false;
    Run cur;
    for (auto& kv : points) {
      if (open && cur.end + 1 == kv.first) {
        cur.end = kv.first;
        cur.point_count += 1;
        cur.item_count += kv.second;
      } else {
        if (open) out.runs.push_back(cur);
        cur = Run{kv.first, kv.first, 1, kv.second};
        open = true;
      }
    }
    if (open) out.runs.push_back(cur);
    return out;
  }
  void reset() {
    points.clear();
    items = 0;
  }
};
// indexed_presence.rs:6-147 — indexed_presence(index): the distinct integer points of a group with their item
// counts; the result offers membership, counts in a range, the active runs and the complement runs of a horizon.
struct IndexedPresence {
  std::map<int64_t, size_t> points;
  size_t items = 0;
  bool contains(int64_t i) const { return points.count(i) != 0; }
  size_t count() const { return points.size(); }
  size_t item_count() const { return items; }
  bool is_empty() const { return points.empty(); }
  static Runs runs_from_counts(const std::map<int64_t, size_t>& pts) {  // runs.rs:179-229
    Runs out;
    out.point_count = pts.size();
    bool open = false;
    Run cur;
    for (auto& kv : pts) {
      out.item_count += kv.second;
      if (open && cur.end + 1 == kv.first) {
        cur.end = kv.first;
        cur.point_count += 1;
        cur.item_count += kv.second;
      } else {
        if (open) out.runs.push_back(cur);
        cur = Run{kv.first, kv.first, 1, kv.second};
        open = true;
      }
    }
    if (open) out.runs.push_back(cur);
    return out;
  }
  Runs runs() const { return runs_from_counts(points); }
  Runs complement_runs(int64_t lo, int64_t hi) const {  // :72-91
    std::map<int64_t, size_t> comp;
    for (int64_t i = lo; i < hi; ++i)
      if (!points.count(i)) comp[i] = 1;
    return runs_from_counts(comp);
  }
  size_t count_in(int64_t lo, int64_t hi) const {  // :93-98
    if (lo >= hi) return 0;
    size_t n = 0;
    for (auto it = points.lower_bound(lo); it != points.end() && it->first < hi; ++it) ++n;
    return n;
  }
  bool any_in(int64_t lo, int64_t hi) const { return count_in(lo, hi) > 0; }
};
struct IndexedPresenceAcc {  // :104-147
  using Value = int64_t;
  using Result = IndexedPresence;
  using Retraction = int64_t;
  IndexedPresence presence;
  Retraction accumulate(Value v) {
    presence.points[v] += 1;
    presence.items += 1;
    return v;
  }
  void retract(Retraction v) {
    auto it = presence.points.find(v);
    if (it == presence.points.end()) return;
    it->second = it->second > 0 ? it->second - 1 : 0;
    presence.items = presence.items > 0 ? presence.items - 1 : 0;
    if (it->second == 0) presence.points.erase(it);
  }
  Result result() const { return presence; }
  void reset() { presence = IndexedPresence{}; }
};
// load_balance.rs:104-240. Result carries `unfairness` (the per-key loads map is not scored).
struct LoadBalanceAcc {
  using Value = std::pair<int64_t, int64_t>;  // (balanced key, metric)
  using Result = int64_t;                      // unfairness
  using Retraction = std::pair<int64_t, int64_t>;
  std::unordered_map<int64_t, size_t> item_counts;
  std::unordered_map<int64_t, int64_t> loads;
  int64_t sum = 0, sq_integral = 0, sq_fraction_num = 0;
  void update_sq(int64_t old_v, int64_t new_v) {  // :143-162
    int64_t term1 = new_v * new_v - old_v * old_v;
    int64_t sum_others = 2 * (sum - old_v);
    int64_t new_sum = sum - old_v + new_v;
    int64_t sum_diff = sum - new_sum;
    int64_t term3 = new_sum * new_sum - sum * sum;
    int64_t term4 = 2 * (old_v * sum - new_v * new_sum);
    sq_integral += term1;
    sq_fraction_num += sum_others * sum_diff + term3 + term4;
  }
  void add_to_metric(int64_t key, int64_t diff) {
    auto it = loads.find(key);
    int64_t old_v = it == loads.end() ? 0 : it->second;
    int64_t new_v = old_v + diff;
    if (old_v != new_v) {
      loads[key] = new_v;
      update_sq(old_v, new_v);
      sum += diff;
    }
  }
  void reset_metric(int64_t key) {
    auto it = loads.find(key);
    if (it == loads.end()) return;
    int64_t old_v = it->second;
    loads.erase(it);
    if (old_v != 0) {
      update_sq(old_v, 0);
      sum -= old_v;
    }
  }
  Retraction accumulate(Value v) {
    if (v.second == 0) return v;
    item_counts[v.first] += 1;
    add_to_metric(v.first, v.second);
    return v;
  }
  void retract(Retraction v) {
    if (v.second == 0) return;
    auto it = item_counts.find(v.first);
    if (it == item_counts.end() || it->second == 0) return;
    it->second -= 1;
    if (it->second == 0) {
      item_counts.erase(it);
      reset_metric(v.first);
    } else {
      add_to_metric(v.first, -v.second);
    }
  }
  // :165-183 — f64 sqrt + round-half-away-from-zero (Rust f64::round == C round()).
  Result result() const {
    size_t n = item_counts.size();
    if (n == 0) return 0;
    double tmp = n == 1 ? (double)sq_fraction_num + (double)sq_integral
                        : ((double)sq_fraction_num / (double)n) + (double)sq_integral;
    return (int64_t)std::round(std::sqrt(tmp));
  }
  void reset() {
    item_counts.clear();
    loads.clear();
    sum = sq_integral = sq_fraction_num = 0;
  }
};

// ---------------------------------------------------------------------------------------------
// group_by(key, collector).penalize(w(key,result))            grouped/state.rs + scorer.rs
template <class S, class A, class K, class Sc, class Acc, class Fi, class KF, class VF, class W,
          class KH = std::hash<K>>
struct GroupedConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> src;
  Impact impact;
  Fi filter;   // (const S&, const A&) -> bool
  KF key_fn;   // (const A&) -> K
  VF value_fn; // (const A&) -> Acc::Value
  W weight;    // (const K&, const Acc::Result&) -> Sc
  struct Group {
    K key;
    Acc acc;
    size_t count = 0;
  };
  std::vector<Group> groups;
  std::unordered_map<K, size_t, KH> group_ids;
  std::unordered_map<size_t, size_t> entity_groups;
  std::unordered_map<size_t, typename Acc::Retraction> entity_retractions;
  std::vector<size_t> changed;
  std::vector<Sc> cached;

  GroupedConstraint(std::string n, Impact i, Source<S, A> s, Fi f, KF kf, VF vf, W w, bool hard)
      : src(s), impact(i), filter(std::move(f)), key_fn(std::move(kf)), value_fn(std::move(vf)),
        weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  Sc evaluate(const S& s) const override {
    std::unordered_map<K, Acc, KH> g;
    for (auto& e : src.extract(s)) {
      if (!src.has(s, e) || !filter(s, e)) continue;
      g[key_fn(e)].accumulate(value_fn(e));
    }
    Sc t = Sc::zero();
    for (auto& kv : g) t = t + signed_weight(impact, weight(kv.first, kv.second.result()));
    return t;
  }
  size_t match_count(const S& s) const override {
    std::unordered_set<K, KH> g;
    for (auto& e : src.extract(s))
      if (src.has(s, e) && filter(s, e)) g.insert(key_fn(e));
    return g.size();
  }
  void mark(size_t g) {
    if (std::find(changed.begin(), changed.end(), g) == changed.end()) changed.push_back(g);
  }
  size_t group_id_for(const K& k) {
    auto it = group_ids.find(k);
    if (it != group_ids.end()) return it->second;
    size_t id = groups.size();
    groups.push_back(Group{k, Acc{}, 0});
    group_ids.emplace(k, id);
    return id;
  }
  void insert_entity(size_t idx, const A& e) {
    size_t g = group_id_for(key_fn(e));
    if (groups[g].count == 0) groups[g].acc.reset();
    entity_retractions[idx] = groups[g].acc.accumulate(value_fn(e));
    groups[g].count += 1;
    entity_groups[idx] = g;
    mark(g);
  }
  void retract_entity(size_t idx) {
    auto ig = entity_groups.find(idx);
    if (ig == entity_groups.end()) return;
    size_t g = ig->second;
    entity_groups.erase(ig);
    auto ir = entity_retractions.find(idx);
    if (ir == entity_retractions.end()) return;
    groups[g].acc.retract(ir->second);
    entity_retractions.erase(ir);
    groups[g].count = groups[g].count > 0 ? groups[g].count - 1 : 0;
    mark(g);
  }
  Sc slot_score(size_t g) const {  // empty group => 0 (state.rs:349-365)
    if (groups[g].count == 0) return Sc::zero();
    return signed_weight(impact, weight(groups[g].key, groups[g].acc.result()));
  }
  Sc replace_cached(size_t slot, Sc sc) {
    while (cached.size() <= slot) cached.push_back(Sc::zero());
    Sc prev = cached[slot];
    cached[slot] = sc;
    return sc - prev;
  }
  Sc refresh_changed() {
    Sc d = Sc::zero();
    for (size_t g : changed) d = d + replace_cached(g, slot_score(g));
    return d;
  }
  Sc initialize(const S& s) override {
    reset();
    auto& es = src.extract(s);
    for (size_t i = 0; i < es.size(); ++i)
      if (src.has(s, es[i]) && filter(s, es[i])) insert_entity(i, es[i]);
    changed.clear();
    Sc t = Sc::zero();
    for (size_t g = 0; g < groups.size(); ++g) {
      Sc sc = slot_score(g);
      replace_cached(g, sc);
      t = t + sc;
    }
    return t;
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    changed.clear();
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    auto& es = src.extract(s);
    if (idx >= es.size()) return Sc::zero();
    if (src.has(s, es[idx]) && filter(s, es[idx])) insert_entity(idx, es[idx]);
    return refresh_changed();
  }
  Sc on_retract(const S&, size_t idx, size_t d) override {
    changed.clear();
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    retract_entity(idx);
    return refresh_changed();
  }
  void reset() override {
    groups.clear();
    group_ids.clear();
    entity_groups.clear();
    entity_retractions.clear();
    changed.clear();
    cached.clear();
  }
};

// ---------------------------------------------------------------------------------------------
// join(A,B).group_by(gk(a,b), collector).complement(T, key_t, default).penalize(w(key,result))
template <class S, class A, class B, class T, class JK, class GK, class Sc, class Acc, class KA, class KB,
          class F, class GF, class VF, class KT, class DF, class W, class JKH = std::hash<JK>,
          class GKH = std::hash<GK>>
struct CrossComplementedGroupedConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> sa;
  Source<S, B> sb;
  Source<S, T> stg;
  Impact impact;
  KA key_a;
  KB key_b;
  F filter;     // (const S&, const A&, const B&, size_t, size_t) -> bool
  GF group_key; // (const A&, const B&) -> GK
  VF value_fn;  // (const A&, const B&) -> Acc::Value
  KT key_t;     // (const T&) -> GK
  DF default_fn;// (const T&) -> Acc::Result
  W weight;     // (const GK&, const Acc::Result&) -> Sc
  bool complemented;  // false => plain cross_grouped (no targets; groups scored directly)

  struct Payload {
    size_t group_id;
    typename Acc::Retraction retraction;
  };
  struct Group {
    GK key;
    Acc acc;
    size_t count = 0;
  };
  JoinRows<JK, JKH, Payload> st;
  std::vector<Group> groups;
  std::unordered_map<GK, size_t, GKH> group_ids;
  std::unordered_map<size_t, std::vector<size_t>> t_by_group;
  std::unordered_map<size_t, size_t> t_index_to_group;
  std::unordered_map<size_t, typename Acc::Result> t_defaults;
  std::vector<size_t> changed_groups, changed_complements;
  std::vector<Sc> cached;

  CrossComplementedGroupedConstraint(std::string n, Impact i, Source<S, A> a, Source<S, B> b, Source<S, T> t,
                                     KA ka, KB kb, F f, GF gf, VF vf, KT kt, DF df, W w, bool hard,
                                     bool with_complement = true)
      : sa(a), sb(b), stg(t), impact(i), key_a(std::move(ka)), key_b(std::move(kb)), filter(std::move(f)),
        group_key(std::move(gf)), value_fn(std::move(vf)), key_t(std::move(kt)), default_fn(std::move(df)),
        weight(std::move(w)), complemented(with_complement) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  // state.rs:173-213 evaluation_state + complemented_scorer.rs:69-78 / scorer.rs:46-55
  Sc evaluate(const S& s) const override {
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    std::unordered_map<JK, std::vector<size_t>, JKH> bidx;
    for (size_t i = 0; i < eb.size(); ++i)
      if (sb.has(s, eb[i])) bidx[key_b(eb[i])].push_back(i);
    std::unordered_map<GK, Acc, GKH> g;
    for (size_t ia = 0; ia < ea.size(); ++ia) {
      if (!sa.has(s, ea[ia])) continue;
      auto it = bidx.find(key_a(ea[ia]));
      if (it == bidx.end()) continue;
      for (size_t ib : it->second) {
        if (!filter(s, ea[ia], eb[ib], ia, ib)) continue;
        g[group_key(ea[ia], eb[ib])].accumulate(value_fn(ea[ia], eb[ib]));
      }
    }
    Sc t = Sc::zero();
    if (!complemented) {
      for (auto& kv : g) t = t + signed_weight(impact, weight(kv.first, kv.second.result()));
      return t;
    }
    for (auto& tg : stg.extract(s)) {
      if (!stg.has(s, tg)) continue;
      GK k = key_t(tg);
      auto it = g.find(k);
      if (it != g.end()) t = t + signed_weight(impact, weight(k, it->second.result()));
      else t = t + signed_weight(impact, weight(k, default_fn(tg)));
    }
    return t;
  }
  size_t match_count(const S& s) const override {
    if (complemented) {
      size_t n = 0;
      for (auto& tg : stg.extract(s)) n += stg.has(s, tg) ? 1 : 0;
      return n;
    }
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    std::unordered_set<GK, GKH> g;
    for (size_t ia = 0; ia < ea.size(); ++ia)
      for (size_t ib = 0; ib < eb.size(); ++ib)
        if (sa.has(s, ea[ia]) && sb.has(s, eb[ib]) && key_a(ea[ia]) == key_b(eb[ib]) &&
            filter(s, ea[ia], eb[ib], ia, ib))
          g.insert(group_key(ea[ia], eb[ib]));
    return g.size();
  }
  void mark(size_t g) {
    if (std::find(changed_groups.begin(), changed_groups.end(), g) == changed_groups.end())
      changed_groups.push_back(g);
  }
  void mark_t(size_t t) {
    if (std::find(changed_complements.begin(), changed_complements.end(), t) == changed_complements.end())
      changed_complements.push_back(t);
  }
  size_t group_id_for(const GK& k) {
    auto it = group_ids.find(k);
    if (it != group_ids.end()) return it->second;
    size_t id = groups.size();
    groups.push_back(Group{k, Acc{}, 0});
    group_ids.emplace(k, id);
    return id;
  }
  void add_match(const S& s, const std::vector<A>& ea, const std::vector<B>& eb, size_t ia, size_t ib) {
    if (st.matches.count({ia, ib})) return;
    if (!sa.has(s, ea[ia]) || !sb.has(s, eb[ib])) return;
    if (!filter(s, ea[ia], eb[ib], ia, ib)) return;
    size_t g = group_id_for(group_key(ea[ia], eb[ib]));
    if (groups[g].count == 0) groups[g].acc.reset();
    auto r = groups[g].acc.accumulate(value_fn(ea[ia], eb[ib]));
    groups[g].count += 1;
    mark(g);
    st.push_row(ia, ib, Payload{g, r});
  }
  void remove_row(size_t row_idx) {
    Payload p = st.remove_row(row_idx);
    if (p.group_id >= groups.size()) return;
    groups[p.group_id].acc.retract(p.retraction);
    groups[p.group_id].count = groups[p.group_id].count > 0 ? groups[p.group_id].count - 1 : 0;
    mark(p.group_id);
  }
  void insert_a(const S& s, const std::vector<A>& ea, const std::vector<B>& eb, size_t ia) {
    if (ia >= ea.size() || !sa.has(s, ea[ia])) return;
    JK k = key_a(ea[ia]);
    auto it = st.b_by_key.find(k);
    std::vector<size_t> bs = it == st.b_by_key.end() ? std::vector<size_t>{} : it->second;
    st.a_by_key[k].push_back(ia);
    st.a_index_to_key[ia] = k;
    for (size_t ib : bs) add_match(s, ea, eb, ia, ib);
  }
  void insert_b(const S& s, const std::vector<A>& ea, const std::vector<B>& eb, size_t ib) {
    if (ib >= eb.size() || !sb.has(s, eb[ib])) return;
    JK k = key_b(eb[ib]);
    auto it = st.a_by_key.find(k);
    std::vector<size_t> as = it == st.a_by_key.end() ? std::vector<size_t>{} : it->second;
    st.b_by_key[k].push_back(ib);
    st.b_index_to_key[ib] = k;
    for (size_t ia : as) add_match(s, ea, eb, ia, ib);
  }
  void retract_a(size_t ia) {
    auto ik = st.a_index_to_key.find(ia);
    if (ik != st.a_index_to_key.end()) {
      st.key_bucket_remove(st.a_by_key, ik->second, ia);
      st.a_index_to_key.erase(ik);
    }
    for (;;) {
      auto it = st.a_to.find(ia);
      if (it == st.a_to.end() || it->second.empty()) break;
      remove_row(it->second.back());
    }
  }
  void retract_b(size_t ib) {
    auto ik = st.b_index_to_key.find(ib);
    if (ik != st.b_index_to_key.end()) {
      st.key_bucket_remove(st.b_by_key, ik->second, ib);
      st.b_index_to_key.erase(ik);
    }
    for (;;) {
      auto it = st.b_to.find(ib);
      if (it == st.b_to.end() || it->second.empty()) break;
      remove_row(it->second.back());
    }
  }
  static void group_bucket_remove(std::unordered_map<size_t, std::vector<size_t>>& m, size_t g, size_t t) {
    auto it = m.find(g);
    if (it == m.end()) return;
    auto& v = it->second;
    auto f = std::find(v.begin(), v.end(), t);
    if (f != v.end()) {
      *f = v.back();
      v.pop_back();
    }
    if (v.empty()) m.erase(it);
  }
  void insert_complement(const S& s, const std::vector<T>& et, size_t t) {
    if (t >= et.size() || !stg.has(s, et[t])) return;
    size_t g = group_id_for(key_t(et[t]));
    t_defaults[t] = default_fn(et[t]);
    auto old = t_index_to_group.find(t);
    if (old != t_index_to_group.end()) {
      group_bucket_remove(t_by_group, old->second, t);
      mark(old->second);
    }
    t_index_to_group[t] = g;
    t_by_group[g].push_back(t);
    mark_t(t);
    mark(g);
  }
  void retract_complement(size_t t) {
    auto it = t_index_to_group.find(t);
    if (it == t_index_to_group.end()) return;
    size_t g = it->second;
    t_index_to_group.erase(it);
    t_defaults.erase(t);
    group_bucket_remove(t_by_group, g, t);
    mark_t(t);
    mark(g);
  }
  // view.rs:24-42 — slot keyed by target index.
  Sc complement_slot_score(size_t t) const {
    auto it = t_index_to_group.find(t);
    if (it == t_index_to_group.end()) return Sc::zero();
    const Group& g = groups[it->second];
    if (g.count > 0) return signed_weight(impact, weight(g.key, g.acc.result()));
    auto d = t_defaults.find(t);
    if (d != t_defaults.end()) return signed_weight(impact, weight(g.key, d->second));
    return Sc::zero();
  }
  Sc group_slot_score(size_t g) const {
    if (groups[g].count == 0) return Sc::zero();
    return signed_weight(impact, weight(groups[g].key, groups[g].acc.result()));
  }
  Sc replace_cached(size_t slot, Sc sc) {
    while (cached.size() <= slot) cached.push_back(Sc::zero());
    Sc prev = cached[slot];
    cached[slot] = sc;
    return sc - prev;
  }
  // view.rs:157-176 — targets of changed groups, then changed complements, each once.
  Sc refresh_changed() {
    Sc d = Sc::zero();
    if (!complemented) {
      for (size_t g : changed_groups) d = d + replace_cached(g, group_slot_score(g));
      return d;
    }
    std::unordered_set<size_t> visited;
    for (size_t g : changed_groups) {
      auto it = t_by_group.find(g);
      if (it == t_by_group.end()) continue;
      for (size_t t : it->second)
        if (visited.insert(t).second) d = d + replace_cached(t, complement_slot_score(t));
    }
    for (size_t t : changed_complements)
      if (visited.insert(t).second) d = d + replace_cached(t, complement_slot_score(t));
    return d;
  }
  Sc initialize(const S& s) override {
    reset();
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    for (size_t i = 0; i < ea.size(); ++i) {
      if (!sa.has(s, ea[i])) continue;
      JK k = key_a(ea[i]);
      st.a_by_key[k].push_back(i);
      st.a_index_to_key.emplace(i, k);
    }
    for (size_t i = 0; i < eb.size(); ++i) {
      if (!sb.has(s, eb[i])) continue;
      JK k = key_b(eb[i]);
      st.b_by_key[k].push_back(i);
      st.b_index_to_key.emplace(i, k);
    }
    if (complemented) {
      auto& et = stg.extract(s);
      for (size_t t = 0; t < et.size(); ++t) insert_complement(s, et, t);
    }
    for (size_t ia = 0; ia < ea.size(); ++ia) {
      if (!sa.has(s, ea[ia])) continue;
      auto it = st.b_by_key.find(key_a(ea[ia]));
      if (it == st.b_by_key.end()) continue;
      std::vector<size_t> bs = it->second;
      for (size_t ib : bs) add_match(s, ea, eb, ia, ib);
    }
    changed_groups.clear();
    changed_complements.clear();
    Sc t = Sc::zero();
    if (complemented) {
      for (auto& kv : t_index_to_group) {
        Sc sc = complement_slot_score(kv.first);
        replace_cached(kv.first, sc);
        t = t + sc;
      }
    } else {
      for (size_t g = 0; g < groups.size(); ++g) {
        Sc sc = group_slot_score(g);
        replace_cached(g, sc);
        t = t + sc;
      }
    }
    return t;
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    changed_groups.clear();
    changed_complements.clear();
    bool ac = sa.change.assert_localizes(d, this->name);
    bool bc = sb.change.assert_localizes(d, this->name);
    bool tc = complemented && stg.change.assert_localizes(d, this->name);
    if (!ac && !bc && !tc) return Sc::zero();
    auto& ea = sa.extract(s);
    auto& eb = sb.extract(s);
    if (ac) insert_a(s, ea, eb, idx);
    if (bc) insert_b(s, ea, eb, idx);
    if (tc) insert_complement(s, stg.extract(s), idx);
    return refresh_changed();
  }
  Sc on_retract(const S&, size_t idx, size_t d) override {
    changed_groups.clear();
    changed_complements.clear();
    bool ac = sa.change.assert_localizes(d, this->name);
    bool bc = sb.change.assert_localizes(d, this->name);
    bool tc = complemented && stg.change.assert_localizes(d, this->name);
    if (!ac && !bc && !tc) return Sc::zero();
    if (ac) retract_a(idx);
    if (bc) retract_b(idx);
    if (tc) retract_complement(idx);
    return refresh_changed();
  }
  void reset() override {
    st.clear();
    groups.clear();
    group_ids.clear();
    t_by_group.clear();
    t_index_to_group.clear();
    t_defaults.clear();
    changed_groups.clear();
    changed_complements.clear();
    cached.clear();
  }
};

// ---------------------------------------------------------------------------------------------
// Projected rows, single source: for_each(src).project(P) with P::MAX_EMITS >= 1
//   stream/projected_stream/source.rs:13-147 (Projection / RowCoordinate / Source),
//   source/single.rs (SingleSource: rows of entity i are (slot 0, i, emit_index)),
//   constraint/projected/uni.rs:61-263 (row_contributions keyed by RowCoordinate, rows_by_owner),
//   constraint/projected/grouped/{state.rs, terminal.rs} (grouped node over projected rows; the
//   group semantics are those of grouped/state.rs: only groups with count > 0 score).
// `project(const A&, std::vector<Out>&)` appends the emitted rows in emit order.
template <class S, class A, class Out, class Sc, class P, class F, class W>
struct ProjectedUniConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> src;
  Impact impact;
  P project;  // (const A&, std::vector<Out>&) -> void
  F filter;   // (const S&, const Out&) -> bool
  W weight;   // (const Out&) -> Sc
  std::unordered_map<std::pair<size_t, size_t>, Sc, PairHash> row_contributions;  // (entity, emit_index)
  std::unordered_map<size_t, std::vector<size_t>> rows_by_owner;
  ProjectedUniConstraint(std::string n, Impact i, Source<S, A> s, P p, F f, W w, bool hard)
      : src(s), impact(i), project(std::move(p)), filter(std::move(f)), weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  Sc evaluate(const S& s) const override {
    Sc t = Sc::zero();
    std::vector<Out> rows;
    for (auto& e : src.extract(s)) {
      rows.clear();
      project(e, rows);
      for (auto& r : rows)
        if (filter(s, r)) t = t + signed_weight(impact, weight(r));
    }
    return t;
  }
  size_t match_count(const S& s) const override {
    size_t n = 0;
    std::vector<Out> rows;
    for (auto& e : src.extract(s)) {
      rows.clear();
      project(e, rows);
      for (auto& r : rows) n += filter(s, r) ? 1 : 0;
    }
    return n;
  }
  Sc insert_rows(const S& s, size_t idx) {
    auto& es = src.extract(s);
    if (idx >= es.size()) return Sc::zero();
    std::vector<Out> rows;
    project(es[idx], rows);
    Sc t = Sc::zero();
    for (size_t j = 0; j < rows.size(); ++j) {
      if (row_contributions.count({idx, j}) || !filter(s, rows[j])) continue;  // uni.rs:103-112
      Sc c = signed_weight(impact, weight(rows[j]));
      row_contributions[{idx, j}] = c;
      rows_by_owner[idx].push_back(j);
      t = t + c;
    }
    return t;
  }
  Sc initialize(const S& s) override {
    reset();
    Sc t = Sc::zero();
    for (size_t i = 0; i < src.extract(s).size(); ++i) t = t + insert_rows(s, i);
    return t;
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    return insert_rows(s, idx);
  }
  Sc on_retract(const S&, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    Sc t = Sc::zero();
    auto it = rows_by_owner.find(idx);
    if (it == rows_by_owner.end()) return t;
    for (size_t j : it->second) {  // uni.rs:114-120
      auto rc = row_contributions.find({idx, j});
      if (rc == row_contributions.end()) continue;
      t = t - rc->second;
      row_contributions.erase(rc);
    }
    rows_by_owner.erase(it);
    return t;
  }
  void reset() override {
    row_contributions.clear();
    rows_by_owner.clear();
  }
};

template <class S, class A, class Out, class K, class Sc, class Acc, class P, class F, class KF, class VF, class W,
          class KH = std::hash<K>>
struct ProjectedGroupedConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> src;
  Impact impact;
  P project;   // (const A&, std::vector<Out>&) -> void
  F filter;    // (const S&, const Out&) -> bool
  KF key_fn;   // (const Out&) -> K
  VF value_fn; // (const Out&) -> Acc::Value
  W weight;    // (const K&, const Acc::Result&) -> Sc
  struct Group {
    K key;
    Acc acc;
    size_t count = 0;
  };
  std::vector<Group> groups;
  std::unordered_map<K, size_t, KH> group_ids;
  struct RowState {
    size_t group;
    typename Acc::Retraction retraction;
  };
  std::unordered_map<std::pair<size_t, size_t>, RowState, PairHash> row_state;  // (entity, emit_index)
  std::unordered_map<size_t, std::vector<size_t>> rows_by_owner;
  std::vector<size_t> changed;
  std::vector<Sc> cached;  // per group slot (grouped/scorer.rs:89-101)
  ProjectedGroupedConstraint(std::string n, Impact i, Source<S, A> s, P p, F f, KF kf, VF vf, W w, bool hard)
      : src(s), impact(i), project(std::move(p)), filter(std::move(f)), key_fn(std::move(kf)),
        value_fn(std::move(vf)), weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  Sc evaluate(const S& s) const override {
    std::unordered_map<K, Acc, KH> g;
    std::vector<Out> rows;
    for (auto& e : src.extract(s)) {
      rows.clear();
      project(e, rows);
      for (auto& r : rows)
        if (filter(s, r)) g[key_fn(r)].accumulate(value_fn(r));
    }
    Sc t = Sc::zero();
    for (auto& kv : g) t = t + signed_weight(impact, weight(kv.first, kv.second.result()));
    return t;
  }
  size_t match_count(const S& s) const override {
    std::unordered_set<K, KH> g;
    std::vector<Out> rows;
    for (auto& e : src.extract(s)) {
      rows.clear();
      project(e, rows);
      for (auto& r : rows)
        if (filter(s, r)) g.insert(key_fn(r));
    }
    return g.size();
  }
  void mark(size_t g) {
    if (std::find(changed.begin(), changed.end(), g) == changed.end()) changed.push_back(g);
  }
  void insert_rows(const S& s, size_t idx) {
    auto& es = src.extract(s);
    if (idx >= es.size()) return;
    std::vector<Out> rows;
    project(es[idx], rows);
    for (size_t j = 0; j < rows.size(); ++j) {
      if (row_state.count({idx, j}) || !filter(s, rows[j])) continue;  // grouped/state.rs insert_row
      K k = key_fn(rows[j]);
      auto it = group_ids.find(k);
      size_t g;
      if (it == group_ids.end()) {
        g = groups.size();
        groups.push_back(Group{k, Acc{}, 0});
        group_ids.emplace(k, g);
      } else {
        g = it->second;
      }
      if (groups[g].count == 0) groups[g].acc.reset();
      auto r = groups[g].acc.accumulate(value_fn(rows[j]));
      groups[g].count += 1;
      row_state[{idx, j}] = RowState{g, r};
      rows_by_owner[idx].push_back(j);
      mark(g);
    }
  }
  void retract_rows(size_t idx) {
    auto it = rows_by_owner.find(idx);
    if (it == rows_by_owner.end()) return;
    for (size_t j : it->second) {
      auto rs = row_state.find({idx, j});
      if (rs == row_state.end()) continue;
      Group& g = groups[rs->second.group];
      g.acc.retract(rs->second.retraction);
      g.count = g.count > 0 ? g.count - 1 : 0;
      mark(rs->second.group);
      row_state.erase(rs);
    }
    rows_by_owner.erase(it);
  }
  Sc refresh() {  // grouped/scorer.rs:89-101,145-152: new - cached for every changed group
    Sc delta = Sc::zero();
    for (size_t g : changed) {
      if (cached.size() <= g) cached.resize(g + 1, Sc::zero());
      Sc now = groups[g].count > 0 ? signed_weight(impact, weight(groups[g].key, groups[g].acc.result())) : Sc::zero();
      delta = delta + (now - cached[g]);
      cached[g] = now;
    }
    changed.clear();
    return delta;
  }
  Sc initialize(const S& s) override {
    reset();
    for (size_t i = 0; i < src.extract(s).size(); ++i) insert_rows(s, i);
    return refresh();
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    insert_rows(s, idx);
    return refresh();
  }
  Sc on_retract(const S&, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    retract_rows(idx);
    return refresh();
  }
  void reset() override {
    groups.clear();
    group_ids.clear();
    row_state.clear();
    rows_by_owner.clear();
    changed.clear();
    cached.clear();
  }
};

// for_each(src).project(P).join(equal(key)).filter(pair).penalize(w(left, right))
//   constraint/projected/bi.rs:251-330,367-392: rows that pass the row filter are self-joined by key; every
//   unordered pair of distinct rows (rows of one entity included) is oriented by RowCoordinate
//   (entity index, then emit index) and scored when the pair filter accepts (left, right).
// Retained state: rows by key; the notification protocol retracts / inserts all rows of one entity.
template <class S, class A, class Out, class K, class Sc, class P, class F, class KF, class PF, class W,
          class KH = std::hash<K>>
struct ProjectedBiConstraint final : IncrementalConstraint<S, Sc> {
  Source<S, A> src;
  Impact impact;
  P project;      // (const A&, std::vector<Out>&) -> void
  F filter;       // (const S&, const Out&) -> bool
  KF key_fn;      // (const Out&) -> K
  PF pair_filter; // (const Out& left, const Out& right) -> bool
  W weight;       // (const Out& left, const Out& right) -> Sc
  struct Row {
    std::pair<size_t, size_t> coord;  // (entity, emit_index)
    Out out;
  };
  std::unordered_map<K, std::vector<Row>, KH> by_key;
  std::unordered_map<size_t, std::vector<std::pair<K, size_t>>> rows_by_owner;  // (key, emit_index)
  ProjectedBiConstraint(std::string n, Impact i, Source<S, A> s, P p, F f, KF kf, PF pf, W w, bool hard)
      : src(s), impact(i), project(std::move(p)), filter(std::move(f)), key_fn(std::move(kf)),
        pair_filter(std::move(pf)), weight(std::move(w)) {
    this->name = std::move(n);
    this->is_hard = hard;
  }
  Sc pair_score(const Row& x, const Row& y) const {
    const Row& l = x.coord <= y.coord ? x : y;
    const Row& r = x.coord <= y.coord ? y : x;
    return pair_filter(l.out, r.out) ? signed_weight(impact, weight(l.out, r.out)) : Sc::zero();
  }
  std::vector<Row> all_rows(const S& s) const {
    std::vector<Row> rows;
    std::vector<Out> tmp;
    auto& es = src.extract(s);
    for (size_t i = 0; i < es.size(); ++i) {
      tmp.clear();
      project(es[i], tmp);
      for (size_t j = 0; j < tmp.size(); ++j)
        if (filter(s, tmp[j])) rows.push_back({{i, j}, tmp[j]});
    }
    return rows;
  }
  Sc evaluate(const S& s) const override {
    auto rows = all_rows(s);
    Sc t = Sc::zero();
    for (size_t a = 0; a < rows.size(); ++a)
      for (size_t b = a + 1; b < rows.size(); ++b)
        if (key_fn(rows[a].out) == key_fn(rows[b].out)) t = t + pair_score(rows[a], rows[b]);
    return t;
  }
  size_t match_count(const S& s) const override {
    auto rows = all_rows(s);
    size_t n = 0;
    for (size_t a = 0; a < rows.size(); ++a)
      for (size_t b = a + 1; b < rows.size(); ++b)
        if (key_fn(rows[a].out) == key_fn(rows[b].out)) {
          const Row& l = rows[a].coord <= rows[b].coord ? rows[a] : rows[b];
          const Row& r = rows[a].coord <= rows[b].coord ? rows[b] : rows[a];
          n += pair_filter(l.out, r.out) ? 1 : 0;
        }
    return n;
  }
  Sc insert_rows(const S& s, size_t idx) {
    auto& es = src.extract(s);
    if (idx >= es.size()) return Sc::zero();
    std::vector<Out> tmp;
    project(es[idx], tmp);
    Sc t = Sc::zero();
    for (size_t j = 0; j < tmp.size(); ++j) {
      if (!filter(s, tmp[j])) continue;
      Row row{{idx, j}, tmp[j]};
      K k = key_fn(tmp[j]);
      auto& bucket = by_key[k];
      for (const Row& other : bucket) t = t + pair_score(row, other);
      bucket.push_back(row);
      rows_by_owner[idx].push_back({k, j});
    }
    return t;
  }
  Sc retract_rows(size_t idx) {
    Sc t = Sc::zero();
    auto it = rows_by_owner.find(idx);
    if (it == rows_by_owner.end()) return t;
    for (auto& kj : it->second) {
      auto& bucket = by_key[kj.first];
      size_t pos = 0;
      for (; pos < bucket.size(); ++pos)
        if (bucket[pos].coord == std::make_pair(idx, kj.second)) break;
      if (pos == bucket.size()) continue;
      Row row = bucket[pos];
      bucket.erase(bucket.begin() + pos);
      for (const Row& other : bucket) t = t - pair_score(row, other);
    }
    rows_by_owner.erase(it);
    return t;
  }
  Sc initialize(const S& s) override {
    reset();
    Sc t = Sc::zero();
    for (size_t i = 0; i < src.extract(s).size(); ++i) t = t + insert_rows(s, i);
    return t;
  }
  Sc on_insert(const S& s, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    return insert_rows(s, idx);
  }
  Sc on_retract(const S&, size_t idx, size_t d) override {
    if (!src.change.assert_localizes(d, this->name)) return Sc::zero();
    return retract_rows(idx);
  }
  void reset() override {
    by_key.clear();
    rows_by_owner.clear();
  }
};

}  // namespace sfo
