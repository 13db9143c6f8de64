// scenegen.h -- scripted-scene generation shared by the device kernel (scenegen.cu) and a host build
// (tests/test_scenegen_host.py compiles this header with g++ and checks it against NumPy / the host generator).
//
// Restates, for one scene = one thread:
//   randomness.py:13-65        derive_seed (sha256 of "<seed>:<part>"), np.random.default_rng(seed)
//                              = SeedSequence -> PCG64 (XSL-RR 128/64), Generator.integers / Generator.uniform
//   scene_generator.py:171-182 scenario sub-seeds, lead_brake.py:18-129, jaywalk.py:29-117 (draw order preserved)
//   scenes/scene.py:61-88, stanley_controller.py:34-49, hero.py:84-86   spawn jitter, start target index
//   control/utils.py:200-269   smooth_and_compute through the Savitzky-Golay operators of engine.py:savgol_operators
//   carlabev.py:108-131, scene.py:142-170  spawn validation + retry loop
// Every random draw, raw route, behaviour parameter and spawn jitter is bit-identical to the reference; the SMOOTHED
// routes agree to ~1e-12 px only: SciPy evaluates the filter edges with a LAPACK least-squares fit whose operation
// order cannot be reproduced instruction for instruction.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD inline
#endif

namespace scenegen {

// ---- sha256 (FIPS 180-4) of a short ASCII message (< 56 bytes: one block) --------------------------------------------
SG_HD uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

SG_HD void sha256_short(const char* msg, int len, uint32_t out[8]) {
  const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
      0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
      0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
      0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
      0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
      0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
      0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t w[64];
  for (int i = 0; i < 16; ++i) w[i] = 0;
  for (int i = 0; i < len; ++i) w[i >> 2] |= (uint32_t)(uint8_t)msg[i] << (24 - 8 * (i & 3));
  w[len >> 2] |= 0x80u << (24 - 8 * (len & 3));
  w[15] = (uint32_t)len * 8u;
  for (int i = 16; i < 64; ++i) {
    uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; ++i) {
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + K[i] + w[i];
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  out[0] = h[0] + a; out[1] = h[1] + b; out[2] = h[2] + c; out[3] = h[3] + d;
  out[4] = h[4] + e; out[5] = h[5] + f; out[6] = h[6] + g; out[7] = h[7] + hh;
}

// derive_seed(base, part) = int(sha256(f"{base}:{part}").hexdigest()[:16], 16) % (2**31 - 1), randomness.py:13-16
SG_HD int64_t derive_seed(int64_t base, const char* part) {
  char msg[48];
  int n = 0;
  char digits[24];
  int nd = 0;
  uint64_t mag = base < 0 ? (uint64_t)(-(base + 1)) + 1u : (uint64_t)base;
  do { digits[nd++] = (char)('0' + mag % 10); mag /= 10; } while (mag);
  if (base < 0) msg[n++] = '-';
  while (nd) msg[n++] = digits[--nd];
  msg[n++] = ':';
  for (const char* p = part; *p; ++p) msg[n++] = *p;
  uint32_t h[8];
  sha256_short(msg, n, h);
  const uint64_t v = ((uint64_t)h[0] << 32) | h[1];
  return (int64_t)(v % 2147483647ull);
}

// ---- np.random.default_rng(seed): SeedSequence (numpy/random/bit_generator.pyx) -> PCG64 (pcg64.h) ---------------------
struct Pcg64 {
  unsigned __int128 state, inc;
  int has_uint32;
  uint32_t uinteger;
};

SG_HD uint32_t ss_hashmix(uint32_t value, uint32_t& hash_const) {
  value ^= hash_const;
  hash_const *= 0x931e8875u;
  value *= hash_const;
  value ^= value >> 16;
  return value;
}
SG_HD uint32_t ss_mix(uint32_t x, uint32_t y) {
  uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
  r ^= r >> 16;
  return r;
}

SG_HD void pcg_step(Pcg64& g) {
  const unsigned __int128 mult = ((unsigned __int128)2549297995355413924ull << 64) | 4865540595714422341ull;
  g.state = g.state * mult + g.inc;
}

// seed >= 0 (SeedSequence rejects negative entropy; derive_seed and scene seeds are non-negative)
SG_HD void pcg_seed(Pcg64& g, uint64_t seed) {
  uint32_t entropy[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int n_ent = entropy[1] ? 2 : 1;  // little-endian 32-bit words of the integer, no trailing zero words
  uint32_t pool[4];
  uint32_t hc = 0x43b0d7e5u;
  for (int i = 0; i < 4; ++i) pool[i] = ss_hashmix(i < n_ent ? entropy[i] : 0u, hc);
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < 4; ++d)
      if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
  // generate_state(4, uint64) = 8 uint32 words, pairs little-endian
  uint32_t st[8];
  uint32_t hb = 0x8b51f9ddu;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= 0x58f38dedu;
    v *= hb;
    v ^= v >> 16;
    st[i] = v;
  }
  const uint64_t w0 = st[0] | ((uint64_t)st[1] << 32), w1 = st[2] | ((uint64_t)st[3] << 32);
  const uint64_t w2 = st[4] | ((uint64_t)st[5] << 32), w3 = st[6] | ((uint64_t)st[7] << 32);
  const unsigned __int128 initstate = ((unsigned __int128)w0 << 64) | w1, initseq = ((unsigned __int128)w2 << 64) | w3;
  g.state = 0;
  g.inc = (initseq << 1) | 1;
  pcg_step(g);
  g.state += initstate;
  pcg_step(g);
  g.has_uint32 = 0;
  g.uinteger = 0;
}

SG_HD uint64_t pcg_next64(Pcg64& g) {
  pcg_step(g);
  const uint64_t hi = (uint64_t)(g.state >> 64), lo = (uint64_t)g.state;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(g.state >> 122);
  return (x >> rot) | (x << ((64 - rot) & 63));
}
SG_HD uint32_t pcg_next32(Pcg64& g) {
  if (g.has_uint32) {
    g.has_uint32 = 0;
    return g.uinteger;
  }
  const uint64_t n = pcg_next64(g);
  g.has_uint32 = 1;
  g.uinteger = (uint32_t)(n >> 32);
  return (uint32_t)n;
}
// Generator.uniform(low, high) = low + (high - low) * next_double, distributions.c:random_uniform
SG_HD double rng_uniform(Pcg64& g, double low, double high) {
  const double u = (double)(pcg_next64(g) >> 11) * (1.0 / 9007199254740992.0);
  return low + (high - low) * u;
}
// Generator.integers(low, high) for ranges below 2^32: Lemire's bounded rejection on 32-bit draws
// (distributions.c:buffered_bounded_lemire_uint32 through random_bounded_uint64_fill)
SG_HD int64_t rng_integers(Pcg64& g, int64_t low, int64_t high) {
  const uint32_t rng = (uint32_t)(high - 1 - low);
  if (rng == 0) return low;
  const uint32_t rng_excl = rng + 1u;
  uint64_t m = (uint64_t)pcg_next32(g) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    const uint32_t threshold = (0xffffffffu - rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)pcg_next32(g) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return low + (int64_t)(m >> 32);
}

// ---- scenario samplers ----------------------------------------------------------------------------------------------------
constexpr int MAX_ACTORS = 3;    // lead_brake level 3
constexpr int MAX_ROUTE = 8;     // pedestrian crossing (np.linspace(..., 8)); vehicles 6 / 7 points
constexpr double MPP = 0.3125;   // 40 / 128
constexpr double WB = 2.9;
enum { KIND_LEAD_BRAKE = 1, KIND_JAYWALK = 2 };
enum { BEH_NONE = 0, BEH_LEAD_BRAKE = 1, BEH_CROSS = 2, BEH_STOP_MID = 3, BEH_STOP_RETURN = 4 };

struct ActorSpec {
  int kind, n, beh;          // 0 vehicle / 1 pedestrian, route points, behaviour id
  double rx[MAX_ROUTE], ry[MAX_ROUTE];
  double speed_mps, beh_p[4];
};
struct Sample {
  double ego_rx[6], ego_ry[6];
  double ego_speed;
  int n_actors;
  ActorSpec actors[MAX_ACTORS];  // vehicles first, then pedestrians (ActorManager order)
};

SG_HD double m2s(double d) { return d / MPP; }  // distance_meters_to_surface, envs/geometry.py:49-50

SG_HD void straight_route(ActorSpec& a, double x, double y0, double step, int n) {
  a.n = n;
  for (int i = 0; i < n; ++i) { a.rx[i] = x; a.ry[i] = y0 - (double)i * step; }
}

// LeadBrakeScenario.sample, lead_brake.py:18-129
SG_HD void sample_lead_brake(int level, Pcg64& g, Sample& s) {
  const int ego_start_y = (int)rng_integers(g, 900, 1000);
  const double lead_gap_m = rng_uniform(g, 4.5, 12.5);
  const double ego_speed = rng_uniform(g, 8.0, 16.0);
  const double lead_speed = ego_speed + rng_uniform(g, -2.0, 2.0);
  const double brake_delay = rng_uniform(g, 1.5, 4.0);
  const double brake_strength = rng_uniform(g, 2.0, 6.0);
  const int x_center = 850;
  const double lane_width = m2s(2.2), ego_step = m2s(6.25), lead_step = m2s(1.56), rear_step = m2s(3.12);
  for (int i = 0; i < 6; ++i) { s.ego_rx[i] = x_center; s.ego_ry[i] = (double)ego_start_y - (double)i * ego_step; }
  s.ego_speed = ego_speed;
  int n = 0;
  {
    ActorSpec& a = s.actors[n++];
    a.kind = 0; a.beh = BEH_LEAD_BRAKE;
    straight_route(a, (double)(x_center - 1), s.ego_ry[0] - m2s(lead_gap_m), lead_step, 6);
    a.speed_mps = fmax(0.0, lead_speed);
    a.beh_p[0] = brake_delay; a.beh_p[1] = brake_strength; a.beh_p[2] = 0.0; a.beh_p[3] = 0.0;
  }
  if (level >= 2) {
    ActorSpec& a = s.actors[n++];
    a.kind = 0; a.beh = BEH_NONE; a.n = 7;
    const double lx = (double)x_center - lane_width;
    for (int i = 0; i < 7; ++i) { a.rx[6 - i] = lx; a.ry[6 - i] = (double)(ego_start_y - i * 20); }  // reversed lists
    a.speed_mps = fmax(0.0, rng_uniform(g, 10.0, 18.0));
    a.beh_p[0] = a.beh_p[1] = a.beh_p[2] = a.beh_p[3] = 0.0;
  }
  if (level >= 3) {
    ActorSpec& a = s.actors[n++];
    const double rear_gap_m = rng_uniform(g, 3.0, 6.0);
    a.kind = 0; a.beh = BEH_LEAD_BRAKE;
    straight_route(a, (double)x_center, s.ego_ry[0] + m2s(rear_gap_m), rear_step, 6);
    a.speed_mps = fmax(0.0, fmax(ego_speed - rng_uniform(g, 1.0, 3.0), 4.0));
    a.beh_p[0] = rng_uniform(g, 2.0, 5.0); a.beh_p[1] = brake_strength; a.beh_p[2] = 0.0; a.beh_p[3] = 0.0;
  }
  s.n_actors = n;
}

// JaywalkScenario.sample, jaywalk.py:29-117
SG_HD void sample_jaywalk(int level, Pcg64& g, Sample& s) {
  const int ego_start_y = (int)rng_integers(g, 900, 1000);
  const double ego_speed = rng_uniform(g, 8.0, 14.0);
  const int ped_x_base = 850;
  const double lane_width = m2s(1.6);
  const double cross_offset_m = rng_uniform(g, -3.0, 3.0);
  const double cross_delay = rng_uniform(g, 1.0, 2.5);
  const double pedestrian_speed = rng_uniform(g, 1.2, 2.2);
  const double ego_step = m2s(6.25), rear_step = m2s(3.12);
  const double yield_duration = rng_uniform(g, 0.8, 1.6);
  for (int i = 0; i < 6; ++i) { s.ego_rx[i] = ped_x_base; s.ego_ry[i] = (double)ego_start_y - (double)i * ego_step; }
  s.ego_speed = ego_speed;
  const double cross_offset = m2s(cross_offset_m);
  const double ped_start_x = ((double)ped_x_base + lane_width) + cross_offset;
  const double ped_end_x = ((double)ped_x_base - lane_width) + cross_offset;
  const double ped_y = s.ego_ry[2] + m2s(rng_uniform(g, -1.0, 1.6));
  ActorSpec ped;
  ped.kind = 1; ped.n = 8;
  {  // np.linspace(start, stop, 8): i * step + start, last element = stop
    const double delta = ped_end_x - ped_start_x, step = delta / 7.0;
    for (int i = 0; i < 8; ++i) {
      double v = step == 0.0 ? ((double)i / 7.0) * delta : (double)i * step;
      ped.rx[i] = v + ped_start_x;
      ped.ry[i] = 1.0 * ped_y;
    }
    ped.rx[7] = ped_end_x;
  }
  ped.speed_mps = fmax(0.0, pedestrian_speed);
  if (level == 1) { ped.beh = BEH_CROSS; ped.beh_p[0] = cross_delay; ped.beh_p[1] = 2.0; ped.beh_p[2] = 0.0; ped.beh_p[3] = 0.0; }
  else if (level == 2) { ped.beh = BEH_STOP_MID; ped.beh_p[0] = cross_delay; ped.beh_p[1] = 0.5; ped.beh_p[2] = -1.0; ped.beh_p[3] = 0.0; }
  else { ped.beh = BEH_STOP_RETURN; ped.beh_p[0] = cross_delay; ped.beh_p[1] = 1.0 / 3.0; ped.beh_p[2] = yield_duration; ped.beh_p[3] = 1.0; }
  int n = 0;
  if (level >= 4) {
    ActorSpec& a = s.actors[n++];
    const double rear_gap_m = rng_uniform(g, 3.0, 6.0);
    a.kind = 0; a.beh = BEH_NONE;
    straight_route(a, (double)ped_x_base, s.ego_ry[0] + m2s(rear_gap_m), rear_step, 6);
    a.speed_mps = fmax(0.0, fmax(ego_speed - rng_uniform(g, 1.0, 3.0), 4.0));
    a.beh_p[0] = a.beh_p[1] = a.beh_p[2] = a.beh_p[3] = 0.0;
  }
  s.actors[n++] = ped;
  s.n_actors = n;
}

// ---- smooth_and_compute (control/utils.py:200-269) for an n-point route without consecutive duplicates, n <= 12 ----
// sg = the n x n Savitzky-Golay operator (row stride SG_STRIDE) of engine.py:savgol_operators
constexpr int SG_STRIDE = 12;
constexpr double PI = 0x1.921fb54442d18p+1, TWO_PI = 0x1.921fb54442d18p+2;

SG_HD void smooth_route(const double* ax, const double* ay, int n, const double* sg, double* cx, double* cy, double* cyaw) {
  for (int i = 0; i < n; ++i) {
    double sx = 0.0, sy = 0.0;
    for (int j = 0; j < n; ++j) {
      sx += sg[i * SG_STRIDE + j] * ax[j];
      sy += sg[i * SG_STRIDE + j] * ay[j];
    }
    cx[i] = sx;
    cy[i] = sy;
  }
  double s[SG_STRIDE];
  s[0] = 0.0;
  for (int i = 1; i < n; ++i) s[i] = s[i - 1] + hypot(cx[i] - cx[i - 1], cy[i] - cy[i - 1]);
  if (s[n - 1] <= 1e-9) {
    for (int i = 0; i < n; ++i) cyaw[i] = 0.0;
    return;
  }
  // np.gradient (second order interior, first order edges) + np.arctan2 + np.unwrap
  double prev = 0.0, corr = 0.0;
  for (int i = 0; i < n; ++i) {
    double gx, gy;
    if (i == 0) {
      gx = (cx[1] - cx[0]) / (s[1] - s[0]);
      gy = (cy[1] - cy[0]) / (s[1] - s[0]);
    } else if (i == n - 1) {
      gx = (cx[n - 1] - cx[n - 2]) / (s[n - 1] - s[n - 2]);
      gy = (cy[n - 1] - cy[n - 2]) / (s[n - 1] - s[n - 2]);
    } else {
      const double d1 = s[i] - s[i - 1], d2 = s[i + 1] - s[i];
      const double a = -(d2) / (d1 * (d1 + d2)), bb = (d2 - d1) / (d1 * d2), c = d1 / (d2 * (d1 + d2));
      gx = a * cx[i - 1] + bb * cx[i] + c * cx[i + 1];
      gy = a * cy[i - 1] + bb * cy[i] + c * cy[i + 1];
    }
    const double p = atan2(gy, gx);
    if (i > 0) {
      const double dd = p - prev;
      double m = fmod(dd + PI, TWO_PI);
      if (m != 0.0 && m < 0.0) m += TWO_PI;
      double ddmod = m - PI;
      if (ddmod == -PI && dd > 0.0) ddmod = PI;
      double pc = ddmod - dd;
      if (fabs(dd) < PI) pc = 0.0;
      corr += pc;
    }
    prev = p;
    cyaw[i] = i > 0 ? p + corr : p;
  }
}

SG_HD int nearest(double x, double y, double yaw, const double* cx, const double* cy, int n) {
  const double fx = x + WB * cos(yaw), fy = y + WB * sin(yaw);
  double best = INFINITY;
  int bi = 0;
  for (int i = 0; i < n; ++i) {  // np.argmin(np.hypot(...)): first minimum
    const double d = hypot(fx - cx[i], fy - cy[i]);
    if (d < best) { best = d; bi = i; }
  }
  return bi;
}

SG_HD int round_half_even_i(double v) { return (int)rint(v); }
SG_HD int rect_left(double c, int pad, int size) { return round_half_even_i((double)pad + c) - (size >> 1); }

// One generated scene, fixed-capacity (what the kernel scatters into the PoolDev arrays)
struct Scene {
  double ego_state0[4], ego_target_speed, len_ego_route;
  int ego_tidx0, num_vehicles, n_actors, attempts;
  double ego_cx[6], ego_cy[6], ego_cyaw[6];
  int32_t rew_rx[6], rew_ry[6];
  double rew_cum[6];
  struct Actor {
    int kind, n, beh, tidx0;
    double state0[4], cruise_px, cruise_mps, beh_p[4];
    double cx[MAX_ROUTE], cy[MAX_ROUTE], cyaw[MAX_ROUTE], raw_x[MAX_ROUTE], raw_y[MAX_ROUTE];
  } actors[MAX_ACTORS];
};

// CarlaBEV.reset for scene in {lead_brake, jaywalk} with scene_seed and level (carlabev.py:96-148): sample, load the
// scene (spawn jitter from the route generator for the ego, from ONE copy of the scenario generator for all actors),
// validate the spawn, retry up to max_attempts with the generators running on.  Returns false when every attempt failed.
// sg_all = [13][12][12] Savitzky-Golay operators; map = class map (0 non-drivable) or null (no validation).
SG_HD bool generate_scene(int kind, int level, int64_t scene_seed, const double* sg_all, const uint8_t* map, int map_w,
                          int map_h, int pad, int max_attempts, Scene& out) {
  Pcg64 scenario_rng, route_rng;
  pcg_seed(scenario_rng, (uint64_t)derive_seed(scene_seed, "scenario"));
  pcg_seed(route_rng, (uint64_t)derive_seed(scene_seed, "route"));
  for (int attempt = 0; attempt < max_attempts; ++attempt) {
    Sample smp;
    if (kind == KIND_LEAD_BRAKE) sample_lead_brake(level, scenario_rng, smp);
    else sample_jaywalk(level, scenario_rng, smp);
    out.attempts = attempt + 1;
    // compute_total_dist_m of the float route (scenes/utils.py) and the int32 reward route (scene.py:192-193)
    double len = 0.0;
    for (int i = 1; i < 6; ++i) len += hypot(smp.ego_rx[i] - smp.ego_rx[i - 1], smp.ego_ry[i] - smp.ego_ry[i - 1]);
    out.len_ego_route = len * MPP;
    double rx[6], ry[6];
    for (int i = 0; i < 6; ++i) {
      out.rew_rx[i] = (int32_t)smp.ego_rx[i];
      out.rew_ry[i] = (int32_t)smp.ego_ry[i];
      rx[i] = out.rew_rx[i];
      ry[i] = out.rew_ry[i];
    }
    double acc = 0.0;
    out.rew_cum[0] = 0.0;
    for (int i = 1; i < 6; ++i) {  // carl_reward_fn.py:20-26
      acc = acc + hypot((double)(out.rew_rx[i] - out.rew_rx[i - 1]), (double)(out.rew_ry[i] - out.rew_ry[i - 1]));
      out.rew_cum[i] = acc;
    }
    // hero: Controller.set_route with jitter from the route generator, then BaseAgent's second stanley_control
    smooth_route(rx, ry, 6, sg_all + 6 * SG_STRIDE * SG_STRIDE, out.ego_cx, out.ego_cy, out.ego_cyaw);
    {
      const double x = out.ego_cx[0] + (double)rng_integers(route_rng, -1, 2);
      const double y = out.ego_cy[0] + (double)rng_integers(route_rng, -1, 2);
      const int t0 = nearest(x, y, 0.0, out.ego_cx, out.ego_cy, 6);
      const double yaw = out.ego_cyaw[t0];
      const int t1 = nearest(x, y, yaw, out.ego_cx, out.ego_cy, 6);
      out.ego_state0[0] = x; out.ego_state0[1] = y; out.ego_state0[2] = yaw; out.ego_state0[3] = smp.ego_speed / MPP;
      out.ego_tidx0 = t0 >= t1 ? t0 : t1;
      out.ego_target_speed = smp.ego_speed / MPP;
    }
    // actors: ActorManager.load deep-copies the dict, so every actor draws its jitter from one COPY of the generator
    Pcg64 copy = scenario_rng;
    out.n_actors = smp.n_actors;
    out.num_vehicles = 0;
    for (int a = 0; a < smp.n_actors; ++a) {
      const ActorSpec& sp = smp.actors[a];
      Scene::Actor& o = out.actors[a];
      o.kind = sp.kind; o.n = sp.n; o.beh = sp.beh;
      if (sp.kind == 0) out.num_vehicles += 1;
      for (int k = 0; k < 4; ++k) o.beh_p[k] = sp.beh_p[k];
      for (int i = 0; i < sp.n; ++i) { o.raw_x[i] = sp.rx[i]; o.raw_y[i] = sp.ry[i]; }
      o.cruise_mps = sp.speed_mps;
      o.cruise_px = sp.speed_mps / MPP;
      smooth_route(sp.rx, sp.ry, sp.n, sg_all + sp.n * SG_STRIDE * SG_STRIDE, o.cx, o.cy, o.cyaw);
      const double x = o.cx[0] + (double)rng_integers(copy, -1, 2);
      const double y = o.cy[0] + (double)rng_integers(copy, -1, 2);
      o.tidx0 = nearest(x, y, 0.0, o.cx, o.cy, sp.n);
      o.state0[0] = x; o.state0[1] = y; o.state0[2] = o.cyaw[o.tidx0]; o.state0[3] = o.cruise_px;
    }
    // Scene.spawn_validation_info, scene.py:142-170
    bool ok = true;
    if (map != nullptr) {
      const double x = out.ego_state0[0], y = out.ego_state0[1];
      int tx = round_half_even_i(x), ty = round_half_even_i(y);
      tx = tx < 0 ? 0 : (tx > map_w - 1 ? map_w - 1 : tx);
      ty = ty < 0 ? 0 : (ty > map_h - 1 ? map_h - 1 : ty);
      if (map[(size_t)ty * map_w + tx] == 0) ok = false;
      // the ego square follows the map scale (hero.py:14-17: 2 / 4 / 8 px at EnvConfig.size 64 / 128 / 256; the map is
      // 8 x size wide), the scripted actors are built with map_size = 128 at every scale (scene_generator.py:65)
      const int msize = map_w >= 8 ? map_w / 8 : 128, mscale = msize <= 1024 ? 1024 / msize : 1;
      const int hw = 32 / mscale > 0 ? 32 / mscale : 1;
      const int hx = rect_left(x, pad, hw), hy = rect_left(y, pad, hw);
      for (int a = 0; a < out.n_actors && ok; ++a) {
        const int size = out.actors[a].kind == 0 ? 4 : 2;
        const int ax = rect_left(out.actors[a].state0[0], pad, size), ay = rect_left(out.actors[a].state0[1], pad, size);
        if (hx < ax + size && hy < ay + size && hx + hw > ax && hy + hw > ay) ok = false;
      }
    }
    if (ok) return true;
  }
  return false;
}

}  // namespace scenegen
