// numpy's legacy global RNG, restated for ONE call pattern: the RANSAC sample draw of the pose estimator.
//
// replaces: the 567 x `np.random.choice(range(n), 4, replace=False)` of PoseEstimator.init_pose_3d
//           (utils/pose.py:139).  numpy (pinned 1.19 in the reference's environment.yml; the legacy stream is
//           frozen by NEP 19, so every later release draws the same numbers) implements that call as
//           RandomState.choice -> permutation(n)[:4] -> shuffle(arange(n)): a Fisher-Yates pass from the end,
//           j = random_interval(i) for i = n-1 .. 1, where random_interval masks a 32-bit MT19937 output
//           down to the bits of i and rejects values above i.  The Python-level call costs ~30 us (17 ms per
//           detection, 90 % of the device pose path); here the same stream is consumed by a tight host loop.
//
// Host code only (no kernel): the state is the caller's `np.random.get_state()` key / position, updated in
// place so that `np.random.set_state()` leaves the global generator exactly where numpy would have left it.
#include <vector>

#include "common.cuh"

namespace {

struct Mt19937 {
  uint32_t* key;   // 624 words
  int pos;
  void refill() {
    constexpr int N = 624, M = 397;
    constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    int kk = 0;
    uint32_t y;
    for (; kk < N - M; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; kk < N - 1; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    pos = 0;
  }
  uint32_t next32() {
    if (pos == 624) refill();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // numpy/random/src/distributions: random_interval for max <= 0xffffffff
  uint32_t interval(uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (next32() & mask)) > max) {}
    return v;
  }
};

}  // namespace

extern "C" int sdfr_np_choice4(uint32_t* mt_key, int32_t* mt_pos, int64_t n, int32_t draws, int32_t* samples_out) {
  SDFR_REQUIRE(mt_key && mt_pos && samples_out, SDFR_E_INVALID, "sdfr_np_choice4: null pointer");
  SDFR_REQUIRE(n >= 4 && n <= 0x7fffffff && draws >= 0, SDFR_E_INVALID, "sdfr_np_choice4: population %lld, draws %d",
               (long long)n, draws);
  SDFR_REQUIRE(*mt_pos >= 0 && *mt_pos <= 624, SDFR_E_INVALID, "sdfr_np_choice4: bad generator position %d", *mt_pos);
  Mt19937 g{mt_key, *mt_pos};
  // Only permutation(n)[:4] is used, so the array is never shuffled.  Pass 1 consumes the stream exactly as the
  // Fisher-Yates loop does (j_i = random_interval(i) for i = n-1 .. 1; a rejected word leaves i where it is, written
  // without a data-dependent branch) and records the j_i.  Pass 2 follows the four output positions back through the
  // transpositions in reverse order of execution (i = 1 .. n-1): position c was fed by j_i when c == i and by i when
  // c == j_i; once i > 3 a tracked position is always below i, so only the second case remains.
  std::vector<uint32_t> jbuf((size_t)n);
  uint32_t tempered[624];
  int tpos = 624;                                        // next tempered word of the current block (624: none tempered yet)
  auto refill_tempered = [&]() {
    if (g.pos == 624) g.refill();
    const int first = g.pos;
    for (int k = first; k < 624; ++k) {
      uint32_t y = g.key[k];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      tempered[k] = y;
    }
    tpos = first;
  };
  for (int32_t d = 0; d < draws; ++d) {
    uint32_t i = (uint32_t)(n - 1);
    while (i >= 1) {
      if (tpos == 624) refill_tempered();              // g.pos == 624 here except before the very first word
      // as many words as this block still has, or until the pass is done
      int k = tpos;
      // the mask only changes when i crosses a power of two: inside such a level the loop-carried work is one
      // compare and one subtract per word
      const uint32_t mask = 0xffffffffu >> __builtin_clz(i);
      const uint32_t level = (mask >> 1) + 1u;           // smallest i with this mask
      for (; k < 624 && i >= level; ++k) {
        const uint32_t v = tempered[k] & mask;
        jbuf[i] = v;                                     // overwritten by the next word when rejected
        i -= v <= i ? 1u : 0u;
      }
      tpos = k;
      g.pos = k;
    }
    uint32_t c[4] = {0u, 1u, 2u, 3u};
    for (uint32_t s = 1; s <= 3 && s < (uint32_t)n; ++s) {
      const uint32_t j = jbuf[s];
      for (int q = 0; q < 4; ++q) c[q] = c[q] == s ? j : (c[q] == j ? s : c[q]);
    }
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    for (uint32_t s = 4; s < (uint32_t)n; ++s) {
      const uint32_t j = jbuf[s];
      c0 = j == c0 ? s : c0;
      c1 = j == c1 ? s : c1;
      c2 = j == c2 ? s : c2;
      c3 = j == c3 ? s : c3;
    }
    int32_t* o = samples_out + (size_t)d * 4;
    o[0] = (int32_t)c0; o[1] = (int32_t)c1; o[2] = (int32_t)c2; o[3] = (int32_t)c3;
  }
  *mt_pos = g.pos;
  return SDFR_OK;
}
