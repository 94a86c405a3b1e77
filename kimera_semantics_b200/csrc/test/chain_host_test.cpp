// CPU check of ksg_chain.cuh (compiled with g++ -ffp-contract=off): the scan formulation against the sequential float32 loop on
// realistic and adversarial chains.   chain_host_test  ->  "chain ok"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ksg_chain.cuh"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 32); }

static float sequential(float s, const std::vector<float>& t) {
  for (float a : t) { volatile float r = s + a; s = r; }
  return s;
}

static int check(float s0, const std::vector<float>& t, const char* what) {
  const float want = sequential(s0, t);
  int bad = 0;
  const int widths[] = {1, 5, 32, 64};
  for (int w : widths) {
    const float got = ksg::chain_sum_reference(s0, t.data(), (long long)t.size(), w);
    if (ksg::chain_bits(got) != ksg::chain_bits(want)) {
      std::printf("MISMATCH %s width %d: got %.9g want %.9g (s0 %.9g, n %zu)\n", what, w, got, want, s0, t.size());
      ++bad;
    }
  }
  return bad;
}

int main() {
  int bad = 0;
  const float lm = std::log(0.9f), ln = std::log(1.0f - 0.9f);
  // 1. realistic: small counts times log(p) / log(1-p), zeros, 92 000 records (one frame of the hottest voxel)
  for (int rep = 0; rep < 6; ++rep) {
    std::vector<float> t(92000);
    for (float& a : t) {
      const float c0 = (float)(rnd() % 4), c1 = (float)(rnd() % 4), c2 = (float)(rnd() % 3);
      a = (c0 * lm + c1 * ln) + c2 * ln;
      if (rnd() % 10 == 0) a = 0.0f;
    }
    bad += check(-0.60205999132f, t, "realistic/first frame");
    bad += check(-2.5e5f, t, "realistic/steady state");
    bad += check(-3.3e7f, t, "realistic/huge");
  }
  // 2. adversarial: random mantissas and exponents, forced ties, subnormals, terms larger than the running value
  for (int rep = 0; rep < 40; ++rep) {
    std::vector<float> t(5000);
    for (float& a : t) {
      uint32_t m = (rnd() & 0x7FFFFFu) | 0x800000u;
      const int e = (int)(rnd() % 36) - 30;
      if (rnd() % 3 == 0) { const int low = 1 + (int)(rnd() % 11); m = ((m >> low) << low) | (1u << (low - 1)); }
      a = -std::ldexp((float)m, e - 23);
      if (rnd() % 20 == 0) a = 0.0f;
      if (rnd() % 100 == 0) a = -1e-42f;
      if (rnd() % 200 == 0) a = -0.0f;
    }
    const float starts[] = {-1e-3f, -0.60205999132f, -777.25f, -3.0e7f, -1.17549435e-38f};
    for (float s0 : starts) bad += check(s0, t, "adversarial");
  }
  // 3. one term repeated: exact integer stepping through many binades
  {
    std::vector<float> t(200000, ln);
    bad += check(-0.60205999132f, t, "constant term");
  }
  if (bad) { std::printf("chain FAILED: %d mismatches\n", bad); return 1; }
  std::printf("chain ok\n");
  return 0;
}
