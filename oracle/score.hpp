// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the
// product path (solverforge_b200/, include/). Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use it.
//
// CPU restatement of the reference's score types.
//   HardSoftScore          solverforge-core/src/score/hard_soft.rs:35-153
//   HardSoftDecimalScore   solverforge-core/src/score/hard_soft_decimal.rs:14-221 (SCALE = 100000)
//   SoftScore              solverforge-core/src/score/soft.rs (single i64 level; used by the
//                          reference's constraint known-answer tests)
//   field-wise + - neg     solverforge-core/src/score/macros.rs:16-48
//   Ord = hard, then soft  hard_soft.rs:130-137
//   hard_score_delta       solverforge-solver/src/phase/hard_delta.rs:11-35
#pragma once
#include <cmath>
#include <cstdint>
#include <string>

namespace sfo {

// Score::multiply / Score::abs (macros.rs:61-71): per level `(x as f64 * m).round() as i64` (round half
// away from zero, saturating cast) and `x.abs()`.
inline int64_t mul_round(int64_t x, double m) {
  const double r = std::round((double)x * m);
  if (r != r) return 0;
  if (r >= 9223372036854775807.0) return INT64_MAX;
  if (r <= -9223372036854775808.0) return INT64_MIN;
  return (int64_t)r;
}
inline int64_t abs_i64(int64_t x) { return x < 0 ? (int64_t)(0 - (uint64_t)x) : x; }

struct SoftScore {
  int64_t v = 0;
  static SoftScore zero() { return {0}; }
  static SoftScore of(int64_t s) { return {s}; }
  SoftScore operator+(SoftScore o) const { return {(int64_t)((uint64_t)v + (uint64_t)o.v)}; }
  SoftScore operator-(SoftScore o) const { return {(int64_t)((uint64_t)v - (uint64_t)o.v)}; }
  SoftScore operator-() const { return {(int64_t)(0 - (uint64_t)v)}; }
  bool operator==(SoftScore o) const { return v == o.v; }
  bool operator!=(SoftScore o) const { return v != o.v; }
  bool operator<(SoftScore o) const { return v < o.v; }
  bool operator>(SoftScore o) const { return v > o.v; }
  bool operator<=(SoftScore o) const { return v <= o.v; }
  bool operator>=(SoftScore o) const { return v >= o.v; }
  static constexpr int levels = 1;
  static bool level_is_hard(int) { return false; }
  int64_t level(int) const { return v; }
  SoftScore abs() const { return {abs_i64(v)}; }
  SoftScore multiply(double m) const { return {mul_round(v, m)}; }
};

// Release-build Rust i64 arithmetic wraps; we wrap explicitly (macros.rs:24-48).
inline int64_t wadd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
inline int64_t wsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }

struct HardSoftScore {
  int64_t hard = 0, soft = 0;
  static HardSoftScore zero() { return {0, 0}; }
  static HardSoftScore of(int64_t h, int64_t s) { return {h, s}; }
  static HardSoftScore of_hard(int64_t h) { return {h, 0}; }
  static HardSoftScore of_soft(int64_t s) { return {0, s}; }
  static HardSoftScore ONE_HARD() { return {1, 0}; }
  static HardSoftScore ONE_SOFT() { return {0, 1}; }
  HardSoftScore operator+(HardSoftScore o) const { return {wadd(hard, o.hard), wadd(soft, o.soft)}; }
  HardSoftScore operator-(HardSoftScore o) const { return {wsub(hard, o.hard), wsub(soft, o.soft)}; }
  HardSoftScore operator-() const { return {wsub(0, hard), wsub(0, soft)}; }
  bool operator==(HardSoftScore o) const { return hard == o.hard && soft == o.soft; }
  bool operator!=(HardSoftScore o) const { return !(*this == o); }
  bool operator<(HardSoftScore o) const { return hard != o.hard ? hard < o.hard : soft < o.soft; }
  bool operator>(HardSoftScore o) const { return o < *this; }
  bool operator<=(HardSoftScore o) const { return !(o < *this); }
  bool operator>=(HardSoftScore o) const { return !(*this < o); }
  bool is_feasible() const { return hard >= 0; }
  static constexpr int levels = 2;
  static bool level_is_hard(int l) { return l == 0; }
  int64_t level(int l) const { return l == 0 ? hard : soft; }
  std::string str() const { return std::to_string(hard) + "hard/" + std::to_string(soft) + "soft"; }
  HardSoftScore abs() const { return {abs_i64(hard), abs_i64(soft)}; }
  HardSoftScore multiply(double m) const { return {mul_round(hard, m), mul_round(soft, m)}; }
};

// Same layout; values pre-scaled by 100000 (hard_soft_decimal.rs:14,45-48,81-86).
struct HardSoftDecimalScore {
  static constexpr int64_t SCALE = 100000;
  int64_t hard = 0, soft = 0;
  static HardSoftDecimalScore zero() { return {0, 0}; }
  static HardSoftDecimalScore of(int64_t h, int64_t s) { return {h * SCALE, s * SCALE}; }
  static HardSoftDecimalScore of_scaled(int64_t h, int64_t s) { return {h, s}; }
  static HardSoftDecimalScore of_hard(int64_t h) { return {h * SCALE, 0}; }
  static HardSoftDecimalScore of_soft(int64_t s) { return {0, s * SCALE}; }
  HardSoftDecimalScore operator+(HardSoftDecimalScore o) const { return {wadd(hard, o.hard), wadd(soft, o.soft)}; }
  HardSoftDecimalScore operator-(HardSoftDecimalScore o) const { return {wsub(hard, o.hard), wsub(soft, o.soft)}; }
  HardSoftDecimalScore operator-() const { return {wsub(0, hard), wsub(0, soft)}; }
  bool operator==(HardSoftDecimalScore o) const { return hard == o.hard && soft == o.soft; }
  bool operator!=(HardSoftDecimalScore o) const { return !(*this == o); }
  bool operator<(HardSoftDecimalScore o) const { return hard != o.hard ? hard < o.hard : soft < o.soft; }
  bool operator>(HardSoftDecimalScore o) const { return o < *this; }
  bool operator<=(HardSoftDecimalScore o) const { return !(o < *this); }
  bool operator>=(HardSoftDecimalScore o) const { return !(*this < o); }
  bool is_feasible() const { return hard >= 0; }
  static constexpr int levels = 2;
  static bool level_is_hard(int l) { return l == 0; }
  int64_t level(int l) const { return l == 0 ? hard : soft; }
  HardSoftDecimalScore abs() const { return {abs_i64(hard), abs_i64(soft)}; }
  HardSoftDecimalScore multiply(double m) const { return {mul_round(hard, m), mul_round(soft, m)}; }
};

// phase/hard_delta.rs:11-35 — first differing Hard-labelled level decides.
enum class HardDelta { None, Improving, Neutral, Worse };
template <class Sc>
inline HardDelta hard_score_delta(Sc previous, Sc candidate) {
  bool saw = false;
  for (int l = 0; l < Sc::levels; ++l) {
    if (!Sc::level_is_hard(l)) continue;
    saw = true;
    if (candidate.level(l) > previous.level(l)) return HardDelta::Improving;
    if (candidate.level(l) < previous.level(l)) return HardDelta::Worse;
  }
  return saw ? HardDelta::Neutral : HardDelta::None;
}

}  // namespace sfo
