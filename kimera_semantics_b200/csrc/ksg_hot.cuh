// ksg_hot.cuh - opt-in pre-pass for the few "hot" voxels of a `merged` frame (ksg_config.hot_voxel_mode = 1, C <= 32).
//
// The voxels next to the camera are crossed by (almost) every bundle: tens of thousands of semantic updates per frame, applied
// strictly in bundle order by ONE warp of the tile kernel today - the critical path of the 2 cm workload (DESIGN.md sections 7, 9).
// A voxel's per-class recurrence  p <- fl(p + a_k)  is a same-sign float32 chain, and a chunk of such a chain reduces to a two-entry
// table independently of the chunks before it (ksg_chain.cuh).  So, before the tile kernel runs:
//   k_hot_find          record runs of one (tile, voxel) with >= kHotThresh records            -> hot segments
//   k_hot_chunk_sums    per (segment, 1024-record chunk): float64 sum of every class column      (parallel over chunks)
//   k_hot_guess         per segment: running float64 prefix -> the binade each chunk will start in, per class
//   k_hot_chunk_tables  per (segment, chunk): rows staged in shared memory, lanes = 32-record sub-chunks, warp scan of the
//                       composed tables, one table per class on the guessed grid                 (parallel over chunks)
//   k_hot_apply         per segment: lanes = classes walk the chunk tables; a chunk whose guess was wrong or that leaves its binade
//                       is re-evaluated by plain sequential addition (rare: same-sign chains cross a binade once per doubling)
// The result - the voxel's finished log-probability row - is handed to k_tile_apply<..., HOTSEM = true>, whose semantic warp then
// skips the record loop for that voxel.  Bit-identical to the sequential order by construction; CPU model and tests:
// tools/exact_float_chain.py (chunked_sum), csrc/test/chain_host_test.cpp.
#pragma once
#include "ksg_chain.cuh"
#include "ksg_kernels.cuh"

namespace ksg {

static constexpr int kHotThresh = 4096;   // records of one voxel in one frame that make it "hot"
static constexpr int kHotChunk = 1024;    // records per chunk (32 lanes x 32 records)
static constexpr int kHotMaxSegs = 2048;  // more hot voxels than this are left to the ordinary path

static constexpr int kHotColStride = kHotChunk + kHotChunk / 32 + 1;   // column-major staging, skewed: conflict-free both ways

__device__ __forceinline__ uint64_t hot_key(uint64_t rec) { return rec >> kRecOrdBits; }   // [tile key | voxel in tile]

__global__ void k_hot_find(DevCfg cfg, MapRef map, const uint64_t* __restrict__ rec, long long n, HotSeg* __restrict__ segs,
                           int* __restrict__ n_segs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t r = rec[i];
  if (r == ~0ull) return;
  const uint64_t key = hot_key(r);
  if (i > 0 && hot_key(rec[i - 1]) == key) return;                        // not the head of its run
  if (i + kHotThresh - 1 >= n || hot_key(rec[i + kHotThresh - 1]) != key) return;
  long long lo = i + kHotThresh, hi = n;                                  // first record of another run
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (hot_key(rec[mid]) == key) lo = mid + 1; else hi = mid; }
  const uint32_t tk = (uint32_t)(r >> 32);
  const int pos = (int)(tk / (uint32_t)cfg.tiles_per_block), tile = (int)(tk % (uint32_t)cfg.tiles_per_block);
  const int slot = map.ht_slot[pos];
  if (slot < 0 || slot >= map.max_blocks) return;                         // pool overflow: flagged elsewhere, ordinary path
  const int v = (int)((r >> kRecOrdBits) & ((1u << kRecVoxBits) - 1u));
  const int s = atomicAdd(n_segs, 1);
  if (s >= kHotMaxSegs) return;
  HotSeg h;
  h.begin = i; h.end = lo;
  h.prior_off = (long long)((uint64_t)slot * cfg.block_stride + (uint64_t)tile * cfg.tile_stride + cfg.head_bytes + (uint64_t)v * cfg.C * 4u);
  h.tile_off = (long long)((uint64_t)slot * cfg.block_stride + (uint64_t)tile * cfg.tile_stride);
  h.first_chunk = 0; h.n_chunks = (int)((lo - i + kHotChunk - 1) / kHotChunk);
  h.vox = v;
  {   // global voxel index of (block, tile, voxel) and its centre, as k_tile_apply derives them
    const I3 bi = unpack_key(map.ht_keys[pos]);
    const int tps = cfg.tiles_per_side, ts = cfg.tile_side_log2, tm = cfg.tile_side - 1;
    const int tx = tile % tps, ty = (tile / tps) % tps, tz = tile / (tps * tps);
    I3 g;
    g.x = bi.x * cfg.vps + tx * cfg.tile_side + (v & tm);
    g.y = bi.y * cfg.vps + ty * cfg.tile_side + ((v >> ts) & tm);
    g.z = bi.z * cfg.vps + tz * cfg.tile_side + (v >> (2 * ts));
    const F3 c = voxel_center(g, cfg.voxel_size);
    h.cx = c.x; h.cy = c.y; h.cz = c.z;
  }
  segs[s] = h;
}

// hot_voxel_mode 2.  A hot voxel in free space sits at (distance, weight) = (+truncation, max_weight), and every record of the frame
// leaves it there: min(max_weight, max_weight + uw) = max_weight for uw >= 0, and the clamped running mean stays at +truncation.
// This kernel CHECKS that, record by record and in parallel (the recurrence is not evaluated, only its fixed point is verified):
// same[seg] stays 1 iff the state is the saturated one and no record moves the distance, the weight or (kColor mode) the colour.
// One CTA per chunk; same[] must be preset to 1.
__global__ void __launch_bounds__(128) k_hot_tsdf_same(DevCfg cfg, Xform T, const HotSeg* __restrict__ segs, const int* __restrict__ chunk_seg,
                                                       const uint8_t* __restrict__ pool, const uint64_t* __restrict__ rec,
                                                       const float4* __restrict__ param, int* __restrict__ same) {
  const int w = blockIdx.x, seg = chunk_seg[w];
  const HotSeg h = segs[seg];
  const float dist0 = ((const float*)(pool + h.tile_off))[h.vox];
  const float wgt0 = ((const float*)(pool + h.tile_off + cfg.plane_f32))[h.vox];
  if (!(dist0 == cfg.tp.trunc && wgt0 == cfg.tp.max_weight)) { if (threadIdx.x == 0) same[seg] = 0; return; }
  const long long b = h.begin + (long long)(w - h.first_chunk) * kHotChunk;
  const long long e = (b + kHotChunk < h.end) ? b + kHotChunk : h.end;
  const uint32_t ord_mask = (1u << kRecOrdBits) - 1u;
  const F3 origin = f3(T.tx, T.ty, T.tz), center = f3(h.cx, h.cy, h.cz);
  const bool keep_blend = cfg.color_mode == 0;
  bool ok = true;
  for (long long i = b + threadIdx.x; i < e; i += blockDim.x) {
    const float4 pr = param[(uint32_t)rec[i] & ord_mask];
    float sdf, uw;
    tsdf_measure(cfg.tp, origin, f3(pr.x, pr.y, pr.z), center, pr.w, sdf, uw);
    if (!(uw >= 0.0f)) { ok = false; break; }                              // weight could drop below the cap (or NaN)
    const float nw = wgt0 + uw;
    const float nd = (sdf * uw + dist0 * wgt0) / nw;
    const float dn = (nd > 0.0f) ? fminf(cfg.tp.trunc, nd) : fmaxf(-cfg.tp.trunc, nd);
    if (__float_as_uint(dn) != __float_as_uint(dist0)) { ok = false; break; }
    if (keep_blend && fabsf(sdf) < cfg.tp.trunc) { ok = false; break; }    // the record would blend the colour
  }
  if (!ok) same[seg] = 0;
}

// One warp per chunk, lanes = classes: float64 column sums of the chunk's (L * freq) rows.
__global__ void k_hot_chunk_sums(int C, const HotSeg* __restrict__ segs, const int* __restrict__ chunk_seg, int n_chunks_total,
                                 const uint64_t* __restrict__ rec, const float* __restrict__ tmp, double* __restrict__ sums) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_chunks_total) return;
  const HotSeg h = segs[chunk_seg[w]];
  const long long b = h.begin + (long long)(w - h.first_chunk) * kHotChunk;
  const long long e = (b + kHotChunk < h.end) ? b + kHotChunk : h.end;
  const uint32_t ord_mask = (1u << kRecOrdBits) - 1u;
  double s = 0.0;
  if (lane < C) {
    long long i = b;
    for (; i + 8 <= e; i += 8) {                                           // eight independent row loads in flight
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(tmp + (size_t)((uint32_t)rec[i + u] & ord_mask) * C + lane);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; i < e; ++i) s += (double)__ldg(tmp + (size_t)((uint32_t)rec[i] & ord_mask) * C + lane);
    sums[(size_t)w * 32 + lane] = s;
  }
}

// One warp per segment, lanes = classes: the grid (exponent of the ulp) every chunk is expected to start on.
__global__ void k_hot_guess(int C, const HotSeg* __restrict__ segs, int n_segs, const uint8_t* __restrict__ pool,
                            const double* __restrict__ sums, int* __restrict__ guess) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_segs || lane >= C) return;
  const HotSeg h = segs[s];
  double run = (double)((const float*)(pool + h.prior_off))[lane];
  for (int k = 0; k < h.n_chunks; ++k) {
    const size_t w = (size_t)(h.first_chunk + k);
    uint32_t m;
    int g;
    chain_decompose((float)run, m, g);
    guess[w * 32 + lane] = g;
    run += sums[w * 32 + lane];
  }
}

// One CTA (4 warps) per chunk.  The chunk's rows are staged in shared memory with coalesced loads; for each class one warp composes
// the 1024 record tables: lane l folds records 32 l .. 32 l + 31 in order, then an inclusive warp scan leaves the chunk's table
// in lane 31.
__global__ void __launch_bounds__(128) k_hot_chunk_tables(int C, const HotSeg* __restrict__ segs, const int* __restrict__ chunk_seg,
                                                          const uint64_t* __restrict__ rec, const float* __restrict__ tmp,
                                                          const int* __restrict__ guess, ChainTable* __restrict__ tables) {
  extern __shared__ float s_cols[];          // [C][kHotColStride]: record r of class c at c * kHotColStride + r + (r >> 5)
  const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const HotSeg h = segs[chunk_seg[w]];
  const long long b = h.begin + (long long)(w - h.first_chunk) * kHotChunk;
  const long long e = (b + kHotChunk < h.end) ? b + kHotChunk : h.end;
  const int cnt = (int)(e - b);
  const uint32_t ord_mask = (1u << kRecOrdBits) - 1u;
  for (int r = warp; r < kHotChunk; r += 4) {                              // one row per warp iteration, lanes = classes
    float v = 0.0f;                                                        // padded records add +0 (identity)
    if (r < cnt && lane < C) v = __ldg(tmp + (size_t)((uint32_t)rec[b + r] & ord_mask) * C + lane);
    if (lane < C) s_cols[lane * kHotColStride + r + (r >> 5)] = v;
  }
  __syncthreads();
  for (int c = warp; c < C; c += 4) {
    const int g = guess[(size_t)w * 32 + c];
    const float* col = s_cols + c * kHotColStride + lane * 33;            // records 32 lane .. 32 lane + 31, skewed by lane
    ChainTable t = chain_record_table(col[0], g);
#pragma unroll 4
    for (int k = 1; k < 32; ++k) t = chain_compose(t, chain_record_table(col[k], g));
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      ChainTable o;
      o.inc[0] = __shfl_up_sync(0xffffffffu, t.inc[0], off);
      o.inc[1] = __shfl_up_sync(0xffffffffu, t.inc[1], off);
      o.par = __shfl_up_sync(0xffffffffu, t.par, off);
      if (lane >= off) t = chain_compose(o, t);
    }
    if (lane == 31) tables[(size_t)w * 32 + c] = t;
  }
}

// One warp per segment, lanes = classes: apply the chunk tables in order; exact fallback for the rare chunk that does not fit.
__global__ void k_hot_apply(int C, const HotSeg* __restrict__ segs, int n_segs, const uint8_t* __restrict__ pool,
                            const uint64_t* __restrict__ rec, const float* __restrict__ tmp, const int* __restrict__ guess,
                            const ChainTable* __restrict__ tables, float* __restrict__ hot_prior, int* __restrict__ n_fallback) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_segs || lane >= C) return;
  const HotSeg h = segs[s];
  const uint32_t ord_mask = (1u << kRecOrdBits) - 1u;
  float p = ((const float*)(pool + h.prior_off))[lane];
  int fallbacks = 0;
  for (int k = 0; k < h.n_chunks; ++k) {
    const size_t w = (size_t)(h.first_chunk + k);
    uint32_t m;
    int g;
    chain_decompose(p, m, g);
    const ChainTable t = tables[w * 32 + lane];
    const uint32_t inc = (m & 1u) ? t.inc[1] : t.inc[0];
    if (p < 0.0f && m >= 0x800000u && g == guess[w * 32 + lane] && m + inc < (1u << 24)) {
      p = chain_make_negative(m + inc, g);
    } else {                                                                // wrong guess, binade crossing, or p >= 0: plain loop
      const long long b = h.begin + (long long)k * kHotChunk;
      const long long e = (b + kHotChunk < h.end) ? b + kHotChunk : h.end;
      long long i = b;
      for (; i + 8 <= e; i += 8) {                                         // loads run ahead, the additions stay in record order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(tmp + (size_t)((uint32_t)rec[i + u] & ord_mask) * C + lane);
#pragma unroll
        for (int u = 0; u < 8; ++u) p += v[u];
      }
      for (; i < e; ++i) p += __ldg(tmp + (size_t)((uint32_t)rec[i] & ord_mask) * C + lane);
      ++fallbacks;
    }
  }
  hot_prior[(size_t)s * 32 + lane] = p;
  if (fallbacks) atomicAdd(n_fallback, fallbacks);
}

}  // namespace ksg
