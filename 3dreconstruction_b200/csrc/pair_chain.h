// pair_chain.h -- WHEN the pairs of a collection may start, and at which rand() offset (host-side bookkeeping of the
// geometric filter, geometric_api.cu; plain C++, tested on the CPU by tests/native/test_pair_chain.cpp).
//
// The reference filters pair after pair and every pair draws from ONE process-wide rand() stream
// (geometric_filter.h:37-102, estimator_acransac.h:166-168), so pair p + 1 starts at
//     offset(p + 1) = offset(p) + sample_size * iterations_run(p),
// and iterations_run(p) is known only once pair p has its first accepted model (or has seen that its whole first phase
// holds none).  The chain below keeps the pairs whose count is still open, oldest first.  Only the front one has a
// definite offset; the others were started ahead on an ASSUMED count of their predecessor:
//   * `iterations` (the whole budget) after two no-model pairs in a row, up to spec_depth pairs deep -- in a large
//     exhaustive collection most pairs are of that kind;
//   * the predecessor's GUESS: it reported a first trigger candidate that still waits for its exact roots; if the
//     candidate is confirmed the count is candidate + 1 + reserve.
// When the front pair's count becomes final the next pair is either confirmed (its offset was right) or refuted: every pair
// still in the chain is restarted (same slot) at corrected offsets.  A speculative pair that finishes before its offset
// is confirmed is HELD in its slot and counted only then.  Whatever the schedule of verdicts, every pair's last start is
// at the reference's offset.
#pragma once
#include <deque>

namespace mvgcuda {
namespace geo {

class PairChain {
 public:
  enum State { kFree = 0, kActive = 1, kHeldDone = 2 };
  struct Slot {
    int state = kFree;
    int pair = -1;
    long long offset = 0;
    int iters_final = -1;  // >= 0 once the pair's iteration count is known
    int iters_guess = -1;  // >= 0 while a first trigger candidate waits for its exact roots
  };

  PairChain(int n_slots, int n_pairs, int sample, int iterations, long long first_offset, int spec_depth = 3, int guess_depth = 2)
      : n_slots_(n_slots), n_pairs_(n_pairs), sample_(sample), iterations_(iterations), chain_offset_(first_offset),
        spec_depth_(spec_depth), guess_depth_(guess_depth) {
    for (int q = 0; q < kMaxSlots; ++q) slots_[q] = Slot();
  }

  static constexpr int kMaxSlots = 64;

  // The next pair that may start now: its slot and offset (the slot becomes kActive and joins the chain).  False when
  // nothing may start at the moment (no free slot, no pair left, or the chain is as deep as the assumptions allow).
  bool admit(int* slot, int* pair, long long* offset) {
    if (next_admit_ >= n_pairs_) return false;
    const int depth = noise_streak_ >= 2 ? spec_depth_ : 0;
    const bool on_guess = !chain_.empty() && static_cast<int>(chain_.size()) <= guess_depth_ && slots_[chain_.back()].iters_guess >= 0;
    if (static_cast<int>(chain_.size()) > depth && !on_guess) return false;
    int sl = -1;
    for (int q = 0; q < n_slots_; ++q)
      if (slots_[q].state == kFree) { sl = q; break; }
    if (sl < 0) return false;
    long long off = chain_offset_;
    if (!chain_.empty()) {
      const Slot& b = slots_[chain_.back()];
      off = b.offset + static_cast<long long>(sample_) * (b.iters_guess >= 0 ? b.iters_guess : iterations_);
    }
    Slot& S = slots_[sl];
    S.state = kActive; S.pair = next_admit_; S.offset = off; S.iters_final = -1; S.iters_guess = -1;
    chain_.push_back(sl);
    *slot = sl; *pair = next_admit_; *offset = off;
    ++next_admit_;
    return true;
  }

  // Pairs with geometry are about (no run of no-model pairs): a pair's first range may be short.
  bool geometry_about() const { return noise_streak_ < 2; }

  // What a verdict says about the pair's iteration count.
  void note_counts(int slot, int iters_final, int iters_guess) {
    Slot& S = slots_[slot];
    if (S.iters_final < 0 && iters_final >= 0) S.iters_final = iters_final;
    if (iters_guess >= 0) S.iters_guess = iters_guess;
  }

  // The pair in `slot` has finished: counted at once when its offset is definite, else held until it is.
  void pair_done(int slot) {
    Slot& S = slots_[slot];
    if (speculative(slot)) {
      S.state = kHeldDone;
    } else {
      S.state = kFree;
      ++done_pairs_;
    }
  }

  // After verdicts: front pairs whose count is final leave the chain and confirm or refute the offset of the next one.
  // restart(slot, pair, offset) -> 0 or an error code, called for every pair that has to start again.
  template <typename Restart>
  int resolve(Restart&& restart) {
    while (!chain_.empty() && slots_[chain_.front()].iters_final >= 0) {
      const Slot& F = slots_[chain_.front()];
      const long long next_off = F.offset + static_cast<long long>(sample_) * F.iters_final;
      noise_streak_ = F.iters_final == iterations_ ? noise_streak_ + 1 : 0;
      chain_.pop_front();
      if (chain_.empty()) { chain_offset_ = next_off; break; }
      if (slots_[chain_.front()].offset == next_off) {
        Slot& N = slots_[chain_.front()];
        if (N.state == kHeldDone) { N.state = kFree; ++done_pairs_; }  // (its count is final too: it leaves on the next turn)
        continue;
      }
      long long off = next_off;  // refuted: everything started after it starts again
      for (size_t c = 0; c < chain_.size(); ++c) {
        Slot& S = slots_[chain_[c]];
        S.state = kActive; S.offset = off; S.iters_final = -1; S.iters_guess = -1;
        const int rc = restart(chain_[c], S.pair, off);
        if (rc) return rc;
        off += static_cast<long long>(sample_) * iterations_;
      }
      ++refuted_;
      break;
    }
    return 0;
  }

  bool speculative(int slot) const {
    if (chain_.empty() || chain_.front() == slot) return false;
    for (size_t c = 1; c < chain_.size(); ++c)
      if (chain_[c] == slot) return true;
    return false;
  }
  bool finished() const { return done_pairs_ >= n_pairs_; }
  int done_pairs() const { return done_pairs_; }
  long long refuted() const { return refuted_; }
  long long next_offset() const { return chain_offset_; }  // after the last pair: the stream position of the whole run
  const Slot& slot(int q) const { return slots_[q]; }

 private:
  int n_slots_, n_pairs_, sample_, iterations_;
  long long chain_offset_;  // offset of the next pair when the chain is empty
  int spec_depth_, guess_depth_;
  Slot slots_[kMaxSlots];
  std::deque<int> chain_;   // slots of the pairs in flight whose count is open, oldest first
  int next_admit_ = 0, done_pairs_ = 0;
  int noise_streak_ = 0;    // pairs in a row that consumed the whole budget
  long long refuted_ = 0;
};

}  // namespace geo
}  // namespace mvgcuda
