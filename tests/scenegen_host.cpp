// Host build of carlabev_env_b200/csrc/scenegen.h for tests/test_scenegen_host.py (g++ -O2 -ffp-contract=off -shared).
#include "../carlabev_env_b200/csrc/scenegen.h"

extern "C" {
int64_t sgh_derive_seed(int64_t base, const char* part) { return scenegen::derive_seed(base, part); }
void sgh_draws(uint64_t seed, int n, double* uni, int64_t* ints, int64_t lo, int64_t hi) {
  scenegen::Pcg64 g;
  scenegen::pcg_seed(g, seed);
  for (int i = 0; i < n; ++i) {  // interleaved like the samplers do: integers buffer half of a 64-bit draw
    ints[i] = scenegen::rng_integers(g, lo, hi);
    uni[i] = scenegen::rng_uniform(g, -2.0, 5.0);
  }
}
int sgh_generate(int kind, int level, int64_t seed, const double* sg, const uint8_t* map, int w, int h, int pad,
                 scenegen::Scene* out) {
  return scenegen::generate_scene(kind, level, seed, sg, map, w, h, pad, 10, *out) ? 1 : 0;
}
int sgh_scene_bytes() { return (int)sizeof(scenegen::Scene); }
}
