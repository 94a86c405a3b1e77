// ksg_eval.cuh — NEXT-4: ground-truth label accuracy of the map against an analytic world, on the device.
// Mirrors SemanticSimulationWorld::generateSemanticSdfFromWorld (kimera_semantics/src/simulation/semantic_simulation_world.cpp:35-97): the
// ground-truth label of a voxel is the label of the world object closest to the voxel centre (voxblox simulation objects: Sphere, Plane,
// Cube; distance functions as in voxblox/simulation/objects.h).  Evaluated over the observed voxels near the surface
// (weight > 0 and |distance| <= band); the reference keeps the ground truth in a second layer, here it is computed on the fly.
// A checkerboard variant (label = 1 + ((floor(x/s) + floor(y/s) + floor(z/s) + object label) mod (C - 1)), the labelling of the synthetic
// benchmark scene) is selected with checker_size > 0; voxels closer than `checker_margin` to a checker boundary are left out.
#pragma once
#include "ksg_kernels.cuh"

namespace ksg {

struct WorldObject { int type; float a[3]; float b[3]; int label; };   // 0 sphere (a centre, b[0] radius), 1 plane (a point, b normal), 2 cube (a centre, b size)

__host__ __device__ __forceinline__ float world_object_distance(const WorldObject& o, F3 p) {
  if (o.type == 0) return norm3(sub(p, f3(o.a[0], o.a[1], o.a[2]))) - o.b[0];
  if (o.type == 1) return dot3(f3(o.b[0], o.b[1], o.b[2]), sub(p, f3(o.a[0], o.a[1], o.a[2])));
  // voxblox Cube::getDistanceToPoint
  float dv[3];
  const float pp[3] = {p.x, p.y, p.z};
  for (int k = 0; k < 3; ++k) dv[k] = fmaxf(fmaxf(o.a[k] - o.b[k] / 2.0f - pp[k], 0.0f), pp[k] - o.a[k] - o.b[k] / 2.0f);
  float d = norm3(f3(dv[0], dv[1], dv[2]));
  if (d < kEps) {   // inside
    for (int k = 0; k < 3; ++k) dv[k] = fmaxf(o.a[k] - o.b[k] / 2.0f - pp[k], pp[k] - o.a[k] - o.b[k] / 2.0f);
    d = fmaxf(dv[0], fmaxf(dv[1], dv[2]));
  }
  return d;
}

// out[0] evaluated voxels, out[1] correct labels, out[2] observed voxels (weight > 0)
__global__ void k_eval_labels(DevCfg cfg, MapRef map, int n_blocks, const WorldObject* __restrict__ objs, int n_objs, float max_dist, float band,
                              float checker_size, float checker_margin, unsigned long long* __restrict__ out) {
  const int per_block = cfg.tiles_per_block;
  const int V = cfg.tile_voxels;
  unsigned long long n_eval = 0, n_ok = 0, n_obs = 0;
  for (long long w = blockIdx.x; w < (long long)n_blocks * per_block; w += gridDim.x) {
    const int slot = (int)(w / per_block), tile = (int)(w % per_block);
    const uint8_t* chunk = map.pool + (uint64_t)slot * cfg.block_stride + (uint64_t)tile * cfg.tile_stride;
    const float* dist = (const float*)chunk;
    const float* wgt = (const float*)(chunk + cfg.plane_f32);
    const uint8_t* label = chunk + 4 * cfg.plane_f32;
    const I3 bi = unpack_key(map.slot_key[slot]);
    const int tps = cfg.tiles_per_side, ts = cfg.tile_side_log2, tm = cfg.tile_side - 1;
    const int tx = tile % tps, ty = (tile / tps) % tps, tz = tile / (tps * tps);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      if (!(wgt[v] > 0.0f)) continue;
      ++n_obs;
      if (!(fabsf(dist[v]) <= band)) continue;
      I3 g;
      g.x = bi.x * cfg.vps + tx * cfg.tile_side + (v & tm);
      g.y = bi.y * cfg.vps + ty * cfg.tile_side + ((v >> ts) & tm);
      g.z = bi.z * cfg.vps + tz * cfg.tile_side + (v >> (2 * ts));
      const F3 c = voxel_center(g, cfg.voxel_size);
      float best = max_dist;
      int gt = 0;
      bool any = false;
      for (int k = 0; k < n_objs; ++k) {
        const float d = world_object_distance(objs[k], c);
        if (d < best) { best = d; gt = objs[k].label; any = true; }   // semantic_simulation_world.cpp:80-86
      }
      if (!any) continue;
      if (checker_size > 0.0f) {
        const float cs[3] = {c.x / checker_size, c.y / checker_size, c.z / checker_size};
        bool near_edge = false;
        int cell = 0;
        for (int k = 0; k < 3; ++k) {
          const float fl = floorf(cs[k]);
          cell += (int)fl;
          const float fr = (cs[k] - fl) * checker_size;
          if (fr < checker_margin || checker_size - fr < checker_margin) near_edge = true;
        }
        if (near_edge) continue;
        int m = (cell + gt) % (cfg.C - 1);
        if (m < 0) m += cfg.C - 1;
        gt = 1 + m;
      }
      ++n_eval;
      if ((int)label[v] == gt) ++n_ok;
    }
  }
  warp_add(&out[0], n_eval);
  warp_add(&out[1], n_ok);
  warp_add(&out[2], n_obs);
}

}  // namespace ksg
