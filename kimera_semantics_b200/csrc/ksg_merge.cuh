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

// ---------------------------------------------------------------------------------------------
// Voxel-granular deltas: the update log of a frame integrated into emptied layers (ksg_fast.cuh: one entry + C floats per updated voxel)
// lists exactly the voxels the delta map holds, with their final state - merging the log equals merging the delta's blocks, at
// (32 + 4 C) bytes per touched voxel instead of whole 16^3 blocks.  Same arithmetic, per voxel, as k_merge_tiles.
// ---------------------------------------------------------------------------------------------
struct MergeCounts { int n[16]; };

// all deltas at once: hash position of every entry's block (inserting it if new); entry e of delta g lives at g * stride + e
__global__ void k_mergev_insert(Counters* cnt, MapRef map, const VoxelUpdate* __restrict__ upd, MergeCounts counts, int n_deltas, long long stride,
                                int* __restrict__ pos_out, int stamp) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_deltas * stride) return;
  const int g = (int)(t / stride);
  const long long e = t - (long long)g * stride;
  if (e >= counts.n[g]) return;
  const VoxelUpdate u = upd[t];
  I3 b; b.x = u.bx; b.y = u.by; b.z = u.bz;
  int pos = -1;
  if (!key_in_range(b)) set_err(cnt, 5);
  else pos = ht_find_or_insert(map, pack_key(b), cnt);
  pos_out[t] = pos;
  if (pos >= 0) {
    const int old = atomicExch(&map.touched_stamp[pos], stamp);
    if (old != stamp) map.touched_list[atomicAdd(&cnt->n_blocks_touched, 1)] = pos;
  }
}

// one delta: one warp per entry (lanes = classes)
__global__ void __launch_bounds__(256) k_mergev_apply(DevCfg cfg, MapRef map, const Luts* __restrict__ luts, const VoxelUpdate* __restrict__ upd,
                                                      const float* __restrict__ pri, const int* __restrict__ pos_of, int n) {
  const int lane = threadIdx.x & 31;
  const int C = cfg.C;
  const float p_init = (float)-0.60205999132;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int pos = pos_of[i];
    if (pos < 0) continue;
    const int slot = map.ht_slot[pos];
    if (slot < 0 || slot >= map.max_blocks) continue;
    const VoxelUpdate u = upd[i];
    const int lin = (int)(u.lin_label & 0xFFFFFFu);
    const int vm = cfg.vps - 1;
    const int lx = lin & vm, ly = (lin / cfg.vps) & vm, lz = lin / (cfg.vps * cfg.vps);
    const int ts = cfg.tile_side_log2, tm = cfg.tile_side - 1;
    const int tile = (lx >> ts) + cfg.tiles_per_side * ((ly >> ts) + cfg.tiles_per_side * (lz >> ts));
    const int v = (lx & tm) + cfg.tile_side * ((ly & tm) + cfg.tile_side * (lz & tm));
    uint8_t* dst = map.pool + (uint64_t)slot * cfg.block_stride + (uint64_t)tile * cfg.tile_stride;
    float* b_d = (float*)dst; float* b_w = (float*)(dst + cfg.plane_f32);
    uint32_t* b_c = (uint32_t*)(dst + 2 * cfg.plane_f32); uint32_t* b_sc = (uint32_t*)(dst + 3 * cfg.plane_f32);
    uint8_t* b_l = dst + 4 * cfg.plane_f32;
    float* b_p = (float*)(dst + cfg.head_bytes);
    const float* a_p = pri + (size_t)i * C;
    const float wa = u.wgt;
    bool diff = false;
    for (int c = lane; c < C; c += 32) diff |= (__float_as_uint(a_p[c]) != __float_as_uint(p_init));
    const bool sem = __ballot_sync(0xffffffffu, diff) != 0u;
    if (!(wa > 0.0f) && !sem) continue;
    uint32_t rgba = b_c[v];
    if (wa > 0.0f) {
      const float wb = b_w[v];
      const float cw = wa + wb;
      if (cw > 0.0f) {
        const float nd = (u.dist * wa + b_d[v] * wb) / cw;
        rgba = blend_two_colors(u.rgba, wa, rgba, wb);
        __syncwarp();
        if (lane == 0) { b_d[v] = nd; b_w[v] = fminf(cw, cfg.tp.max_weight); }
      }
    }
    float best = -3.402823466e38f;
    int bi_c = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float np = b_p[(size_t)v * C + c] + (a_p[c] - p_init);
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

}  // namespace ksg
