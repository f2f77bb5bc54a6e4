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
  std::vector<int32_t> perm((size_t)n);
  for (int32_t d = 0; d < draws; ++d) {
    for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = (int32_t)i;          // permutation(n): arange, then shuffle
    // Fisher-Yates from the end with numpy's masked rejection, written without a data-dependent branch: a rejected
    // candidate swaps position i with itself and leaves i where it is (the rejection branch of the textbook loop
    // mispredicts on a third of the draws and was 2/3 of the time).
    uint32_t i = (uint32_t)(n - 1);
    while (i >= 1) {
      uint32_t mask = i;
      mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
      const uint32_t v = g.next32() & mask;
      const uint32_t accept = v <= i ? 1u : 0u;
      const uint32_t j = accept ? v : i;
      const int32_t t = perm[i];
      perm[i] = perm[j];
      perm[j] = t;
      i -= accept;
    }
    for (int k = 0; k < 4; ++k) samples_out[(size_t)d * 4 + k] = perm[(size_t)k];   // [:size]
  }
  *mt_pos = g.pos;
  return SDFR_OK;
}
