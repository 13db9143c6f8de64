// scenegen.cu -- device-side generation of the scripted scenarios: one thread per scene
// runs scenegen::generate_scene (scenegen.h: sha256 sub-seeds -> SeedSequence -> PCG64 -> the draws of
// lead_brake.py:18-129 / jaywalk.py:29-117 -> spawn jitter -> route smoothing -> spawn validation with retries) and
// scatters the result into the flat pool arrays the step kernels read (engine.h:PoolDev).  Compiled with -fmad=false.
#include "engine.h"
#include "scenegen.h"

namespace {

__global__ void __launch_bounds__(64)
k_gen_scripted(PoolDev pool, const uint8_t* __restrict__ kinds, const int32_t* __restrict__ levels,
               const long long* __restrict__ seeds, const uint8_t* __restrict__ map, int map_w, int map_h, int pad,
               int32_t* __restrict__ attempts) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= pool.n_scenes) return;
  scenegen::Scene sc;
  const bool ok = scenegen::generate_scene(kinds[s], levels[s], seeds[s], pool.sg_mat, map, map_w, map_h, pad, 10, sc);
  attempts[s] = ok ? sc.attempts : -1;
  for (int k = 0; k < 4; ++k) pool.ego_state0[(size_t)s * 4 + k] = sc.ego_state0[k];
  pool.ego_target_speed[s] = sc.ego_target_speed;
  pool.len_ego_route[s] = sc.len_ego_route;
  pool.ego_tidx0[s] = sc.ego_tidx0;
  pool.num_vehicles[s] = sc.num_vehicles;
  const int e0 = pool.ego_off[s], q0 = pool.rew_off[s];
  for (int i = 0; i < 6; ++i) {
    pool.ego_cx[e0 + i] = sc.ego_cx[i]; pool.ego_cy[e0 + i] = sc.ego_cy[i]; pool.ego_cyaw[e0 + i] = sc.ego_cyaw[i];
    pool.rew_rx[q0 + i] = sc.rew_rx[i]; pool.rew_ry[q0 + i] = sc.rew_ry[i]; pool.rew_cum[q0 + i] = sc.rew_cum[i];
  }
  const int a0 = pool.actor_off[s];
  for (int a = 0; a < sc.n_actors; ++a) {
    const scenegen::Scene::Actor& A = sc.actors[a];
    const int ga = a0 + a;
    pool.act_kind[ga] = (uint8_t)A.kind;
    pool.act_beh[ga] = (uint8_t)A.beh;
    pool.act_tidx0[ga] = A.tidx0;
    pool.act_cruise_px[ga] = A.cruise_px;
    pool.act_cruise_mps[ga] = A.cruise_mps;
    for (int k = 0; k < 4; ++k) {
      pool.act_state0[(size_t)ga * 4 + k] = A.state0[k];
      pool.act_beh_p[(size_t)ga * 4 + k] = A.beh_p[k];
    }
    const int r0 = pool.act_route_off[ga], w0 = pool.act_raw_off[ga];
    for (int i = 0; i < A.n; ++i) {
      pool.act_cx[r0 + i] = A.cx[i]; pool.act_cy[r0 + i] = A.cy[i]; pool.act_cyaw[r0 + i] = A.cyaw[i];
      pool.act_raw_x[w0 + i] = A.raw_x[i]; pool.act_raw_y[w0 + i] = A.raw_y[i];
    }
  }
}

}  // namespace

void cbev_launch_generate(cbev_engine* e, const PoolDev& pool, const uint8_t* kinds, const int32_t* levels,
                          const long long* seeds, int32_t* attempts, cudaStream_t s) {
  const int blocks = (pool.n_scenes + 63) / 64;
  k_gen_scripted<<<blocks, 64, 0, s>>>(pool, kinds, levels, seeds, e->map, e->map_w, e->map_h, e->pad, attempts);
  e->launches += 1;
}
