// ksg_merge.cuh — frame-per-GPU batch mode (BASELINE config 4): merge of one integrator's map ("delta": one frame
// integrated into an empty map on another GPU) into this map, voxel by voxel.
//
// Semantics of the batch mode (the oracle side is tests/delta_merge_ref.py, same arithmetic in numpy float32):
//   a voxel of the delta counts iff its TSDF weight is > 0 or any of its log-probabilities differs from the initial value;
//   TSDF   : voxblox mergeVoxelAIntoVoxelB (the rule voxblox's mergeLayerAintoLayerB applies to sub-maps):
//              w = w_a + w_b ;  d = (d_a * w_a + d_b * w_b) / w ;  colour = blendTwoColors(c_a, w_a, c_b, w_b) ;  weight = min(w, max_weight)
//   label  : log-probabilities add: p_b[c] += (p_a[c] - p_init)  (what semantic_integrator_base.cpp:283-314 would have added to p_b had
//            the delta's observations been integrated into b directly), then arg-max (base.cpp:352-367) and the colour hand-off
//            (base.cpp:370-380, 172-191) as after an ordinary update.
// Deltas are merged in frame order, so the result is a deterministic function of the batch.
#pragma once
#include "ksg_kernels.cuh"

namespace ksg {

__global__ void k_merge_insert(Counters* cnt, MapRef map, const uint64_t* __restrict__ keys, int n, int* __restrict__ pos_out, int stamp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = ht_find_or_insert(map, keys[i], cnt);
  pos_out[i] = pos;
  if (pos >= 0) {   // updated() bookkeeping: every merged block counts as touched by this call
    const int old = atomicExch(&map.touched_stamp[pos], stamp);
    if (old != stamp) map.touched_list[atomicAdd(&cnt->n_blocks_touched, 1)] = pos;
  }
}

// one CTA per (delta block, tile), one warp per voxel (lanes = classes)
__global__ void __launch_bounds__(256) k_merge_tiles(DevCfg cfg, Counters* cnt, MapRef map, const Luts* __restrict__ luts, const int* __restrict__ pos_of,
                                                     const uint8_t* __restrict__ src_pool, int n_blocks) {
  const int per_block = cfg.tiles_per_block;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = cfg.C, V = cfg.tile_voxels;
  const float p_init = (float)-0.60205999132;
  for (long long w = blockIdx.x; w < (long long)n_blocks * per_block; w += gridDim.x) {
    const int bi = (int)(w / per_block), tile = (int)(w % per_block);
    const int pos = pos_of[bi];
    if (pos < 0) continue;
    const int slot = map.ht_slot[pos];
    if (slot < 0 || slot >= map.max_blocks) continue;
    uint8_t* dst = map.pool + (uint64_t)slot * cfg.block_stride + (uint64_t)tile * cfg.tile_stride;
    const uint8_t* src = src_pool + (uint64_t)bi * cfg.block_stride + (uint64_t)tile * cfg.tile_stride;
    const float* a_d = (const float*)src; const float* a_w = (const float*)(src + cfg.plane_f32);
    const uint32_t* a_c = (const uint32_t*)(src + 2 * cfg.plane_f32);
    const float* a_p = (const float*)(src + cfg.head_bytes);
    float* b_d = (float*)dst; float* b_w = (float*)(dst + cfg.plane_f32);
    uint32_t* b_c = (uint32_t*)(dst + 2 * cfg.plane_f32); uint32_t* b_sc = (uint32_t*)(dst + 3 * cfg.plane_f32);
    uint8_t* b_l = dst + 4 * cfg.plane_f32;
    float* b_p = (float*)(dst + cfg.head_bytes);
    for (int v = warp; v < V; v += nwarps) {
      const float wa = a_w[v];
      bool diff = false;
      for (int c = lane; c < C; c += 32) diff |= (__float_as_uint(a_p[(size_t)v * C + c]) != __float_as_uint(p_init));
      const bool sem = __ballot_sync(0xffffffffu, diff) != 0u;
      if (!(wa > 0.0f) && !sem) continue;
      uint32_t rgba = b_c[v];
      if (wa > 0.0f) {
        const float wb = b_w[v];
        const float cw = wa + wb;
        if (cw > 0.0f) {
          const float nd = (a_d[v] * wa + b_d[v] * wb) / cw;
          rgba = blend_two_colors(a_c[v], wa, rgba, wb);
          if (lane == 0) { b_d[v] = nd; b_w[v] = fminf(cw, cfg.tp.max_weight); }
        }
      }
      // log-probabilities: lanes = classes; arg-max, first maximum wins
      float best = -3.402823466e38f;
      int bi_c = 0x7fffffff;
      for (int c = lane; c < C; c += 32) {
        const float np = b_p[(size_t)v * C + c] + (a_p[(size_t)v * C + c] - p_init);
        b_p[(size_t)v * C + c] = np;
        if (np > best || bi_c == 0x7fffffff) { best = np; bi_c = c; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi_c, o);
        if (oi != 0x7fffffff && (bi_c == 0x7fffffff || ob > best || (ob == best && oi < bi_c))) { best = ob; bi_c = oi; }
      }
      if (lane == 0) {
        const uint32_t sc = luts->label_rgba[bi_c];
        b_l[v] = (uint8_t)bi_c;
        b_sc[v] = sc;
        if (cfg.color_mode == 1) rgba = sc;
        else if (cfg.color_mode == 2) rgba = rainbow_color_map((double)expf(best));
        b_c[v] = rgba;
      }
    }
  }
}

}  // namespace ksg
