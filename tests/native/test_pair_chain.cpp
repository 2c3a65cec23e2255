// Host-only model test of csrc/pair_chain.h (the order and the rand() offsets at which the geometric filter starts its
// pairs): random collections of no-model pairs (whole budget) and pairs with geometry (early trigger, right or wrong
// guesses), up to 16 pairs in flight whose verdicts arrive in RANDOM order, runs started at a wrong offset reporting
// arbitrary counts.  Whatever the schedule: every pair is counted once, its last start is at the reference's offset
// (offset0 + sample * sum of the true counts before it), held pairs are released only with a confirmed offset, and the
// final stream position is the sequential one.  TEST INFRASTRUCTURE (no GPU, no reference needed).
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>

#include "../../3dreconstruction_b200/csrc/pair_chain.h"

using mvgcuda::geo::PairChain;

struct Run {            // what the launches of one slot will report, verdict after verdict
  int pair = -1;
  long long offset = 0;
  bool truthful = false;
  std::vector<int> fin, guess, done;  // per verdict: iters_final (or -1), iters_guess (or -1), finished?
  size_t next = 0;
  bool in_flight = false, discard = false;
};

int main() {
  std::mt19937 rng(2024);
  long long cases = 0, refuted_total = 0, held_total = 0, bad = 0;
  for (int trial = 0; trial < 4000 && !bad; ++trial) {
    const int sample = trial & 1 ? 7 : 4, iterations = 4096, reserve = iterations / 10;
    const int n_pairs = 1 + rng() % 60, n_slots = 1 + rng() % 16;
    const long long offset0 = rng() % 1000;
    const double geo_share = (trial % 5) * 0.2;
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::vector<int> count(n_pairs);
    std::vector<bool> geo(n_pairs);
    for (int p = 0; p < n_pairs; ++p) {
      geo[p] = u(rng) < geo_share;
      count[p] = geo[p] ? (int)(rng() % 300) + 1 + reserve : iterations;
    }
    std::vector<long long> off_true(n_pairs + 1, offset0);
    for (int p = 0; p < n_pairs; ++p) off_true[p + 1] = off_true[p] + (long long)sample * count[p];

    PairChain chain(n_slots, n_pairs, sample, iterations, offset0);
    std::vector<Run> run(n_slots);
    std::vector<int> done_count(n_pairs, 0);
    std::vector<long long> last_start(n_pairs, -1);
    auto plan = [&](int sl, int pair, long long off) {
      Run& R = run[sl];
      if (R.in_flight) R.discard = true;      // as geometric_api.cu: the verdict of the launches in flight is dropped
      R.pair = pair; R.offset = off; R.truthful = off == off_true[pair]; R.next = 0;
      R.fin.clear(); R.guess.clear(); R.done.clear();
      last_start[pair] = off;
      const int rounds = 2 + rng() % 8;
      const int c = R.truthful ? count[pair] : (rng() % 3 ? iterations : (int)(rng() % 400) + 1 + reserve);
      const int at_final = rng() % rounds;                    // the verdict that carries the final count
      const bool guesses = (R.truthful ? geo[pair] : (rng() & 1)) && at_final > 0;
      for (int r = 0; r < rounds; ++r) {
        R.fin.push_back(r >= at_final ? c : -1);
        // a guess comes with the verdict before the final one; mostly right, sometimes not
        R.guess.push_back(guesses && r == at_final - 1 ? (u(rng) < 0.85 ? c : c + 1 + (int)(rng() % 50)) : -1);
        R.done.push_back(r == rounds - 1);
      }
      return 0;
    };
    long long steps = 0;
    while (!chain.finished()) {
      if (++steps > 200000) { std::printf("trial %d: no progress (deadlock)\n", trial); ++bad; break; }
      int sl, pair; long long off;
      while (chain.admit(&sl, &pair, &off)) plan(sl, pair, off);
      // launch everything that is active and idle
      std::vector<int> flying;
      for (int q = 0; q < n_slots; ++q) {
        if (chain.slot(q).state == PairChain::kActive && !run[q].in_flight) run[q].in_flight = true;
        if (run[q].in_flight) flying.push_back(q);
      }
      if (flying.empty()) { std::printf("trial %d: nothing in flight\n", trial); ++bad; break; }
      // a random non-empty subset of the launches in flight completes, in random order
      std::shuffle(flying.begin(), flying.end(), rng);
      const size_t k = 1 + rng() % flying.size();
      std::vector<int> held_before;
      for (int q = 0; q < n_slots; ++q) if (chain.slot(q).state == PairChain::kHeldDone) held_before.push_back(q);
      for (size_t f = 0; f < k; ++f) {
        const int q = flying[f];
        Run& R = run[q];
        R.in_flight = false;
        if (R.discard) { R.discard = false; continue; }
        const size_t v = R.next++;
        chain.note_counts(q, R.fin[v], R.done[v] ? -1 : R.guess[v]);
        if (R.done[v]) {
          const bool spec = chain.speculative(q);
          chain.pair_done(q);
          if (!spec) {
            ++done_count[R.pair];
            if (R.offset != off_true[R.pair]) { std::printf("trial %d: pair %d counted from a wrong offset\n", trial, R.pair); ++bad; }
          } else {
            ++held_total;
          }
        }
      }
      const long long before = chain.refuted();
      chain.resolve(plan);
      refuted_total += chain.refuted() - before;
      for (int q : held_before) {
        if (chain.slot(q).state == PairChain::kFree) {   // released: its offset was confirmed
          ++done_count[run[q].pair];
          if (run[q].offset != off_true[run[q].pair]) { std::printf("trial %d: held pair %d released with a wrong offset\n", trial, run[q].pair); ++bad; }
        }
      }
      // (a pair held and released within the same resolve() -- done in this very step -- is accounted for here)
      for (size_t f = 0; f < k; ++f) {
        const int q = flying[f];
        bool was_held_before = false;
        for (int h : held_before) was_held_before |= h == q;
        if (!was_held_before && run[q].next > 0 && run[q].next == run[q].done.size() && run[q].done.back() && chain.slot(q).state == PairChain::kFree &&
            done_count[run[q].pair] == 0 && last_start[run[q].pair] == run[q].offset) {
          ++done_count[run[q].pair];
          if (run[q].offset != off_true[run[q].pair]) { std::printf("trial %d: pair %d released with a wrong offset\n", trial, run[q].pair); ++bad; }
        }
      }
    }
    if (bad) break;
    for (int p = 0; p < n_pairs; ++p) {
      if (done_count[p] != 1) { std::printf("trial %d: pair %d counted %d times\n", trial, p, done_count[p]); ++bad; break; }
      if (last_start[p] != off_true[p]) { std::printf("trial %d: pair %d last started at %lld, reference offset %lld\n", trial, p, last_start[p], off_true[p]); ++bad; break; }
    }
    if (chain.next_offset() != off_true[n_pairs]) { std::printf("trial %d: final offset %lld, sequential %lld\n", trial, chain.next_offset(), off_true[n_pairs]); ++bad; }
    if (chain.done_pairs() != n_pairs) { std::printf("trial %d: %d of %d pairs done\n", trial, chain.done_pairs(), n_pairs); ++bad; }
    ++cases;
  }
  std::printf("%lld collections, %lld refuted starts, %lld pairs held until their offset was confirmed: %s\n", cases, refuted_total, held_total,
              bad ? "FAILED" : "every pair counted once, from the reference's offset");
  std::printf(bad ? "PAIR CHAIN FAILED\n" : "PAIR CHAIN OK\n");
  return bad ? 1 : 0;
}
