// ks_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A dependency-free, scalar C++17 restatement of the Kimera-Semantics semantic TSDF integrators
// (`fast` and `merged`) and of the voxblox primitives they call.  It exists to CHECK the CUDA
// path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs) and is
// never linked, imported or executed by the product path in kimera_semantics_b200/.
//
// PARITY STATUS: pinned for the in-tree half, "parity unpinned" for the voxblox half.
//   * In-tree half (kimera_semantics/src/*.cpp, followed line by line here): pinned against the reference's OWN
//     translation units.  `make -C oracle ref` compiles semantic_tsdf_integrator_{fast,merged}.cpp,
//     semantic_integrator_base.cpp, color.cpp and csv_iterator.cpp where they lie under /root/reference against
//     stand-in dependency headers (oracle/ref_stubs/) into oracle/_ref/libks_ref_hybrid.so;
//     tests/test_oracle_vs_ref_hybrid.py requires this oracle to equal that library bit for bit on 32 seeded
//     sequences (every Config / SemanticConfig switch of the path), and the digests are committed as
//     tests/golden/ref_hybrid_golden.json for boxes without /root/reference.
//   * voxblox half (RayCaster, updateTsdfVoxel, ApproxHashSet, ThreadSafeIndex, bundleRays, Layer/Block, hashes,
//     minkindr transform): the reference ships no tests, golden vectors or fixtures, cannot be
//     built as it stands (catkin, voxblox, minkindr, Eigen, glog absent) and does not vendor or pin voxblox
//     (install/kimera_semantics_https.rosinstall:34-36).  It is restated from the published upstream ethz-asl/voxblox
//     sources as summarised in SURVEY.md Appendix A -- twice, independently (here and in oracle/ref_stubs/voxblox), plus
//     a third time in numpy for the control flow (tests/test_oracle_crosscheck.py) -- and pinned only by this repo's
//     own known-answer tests (tests/test_oracle_kat.py).  Agreement of restatements is not agreement with voxblox.
//
// Build: parity build `-O2 -ffp-contract=off` (no FMA contraction, so every float expression
// rounds exactly as written); timing build `-O3 -march=native -ffp-contract=off`.
//
// Citations: fast.cpp / merged.cpp / base.cpp / base.h / color.cpp are the files under
// /root/reference/kimera_semantics/{src,include/kimera_semantics}/ (see SURVEY.md header table);
// "A.n" = SURVEY.md Appendix A section n (voxblox behaviour).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/ksg.h"

namespace kso {

// ---------------------------------------------------------------------------------------------
// A.0 types
// ---------------------------------------------------------------------------------------------
struct Point { float x, y, z; };
struct GIdx { int64_t x, y, z; bool operator==(const GIdx& o) const { return x == o.x && y == o.y && z == o.z; } };
struct BIdx { int32_t x, y, z; bool operator==(const BIdx& o) const { return x == o.x && y == o.y && z == o.z; } };
struct Color { uint8_t r = 0, g = 0, b = 0, a = 0; };
static constexpr float kEpsilon = 1e-6f;  // kEpsilon = kFloatEpsilon = kCoordinateEpsilon (A.0)

static inline Point operator+(Point a, Point b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline Point operator-(Point a, Point b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline Point operator*(Point a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline Point operator/(Point a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// Eigen fixed-size 3-vector reductions are unrolled left to right: (x*x + y*y) + z*z.
static inline float dot(Point a, Point b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float squaredNorm(Point a) { return dot(a, a); }
static inline float norm(Point a) { return std::sqrt(squaredNorm(a)); }
// Eigen normalized(): returns the vector unchanged when squaredNorm == 0.
static inline Point normalized(Point a) {
  const float n2 = squaredNorm(a);
  if (n2 > 0.0f) return a / std::sqrt(n2);
  return a;
}
static inline Point cross(Point a, Point b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// A.8 minkindr QuatTransformation<float>: T*p = rotate(q,p) + t, Eigen quaternion-vector product
//   uv = q.vec x p; uv += uv; p + q.w*uv + q.vec x uv
struct Transformation {
  float qw, qx, qy, qz;
  Point t;
  Point getPosition() const { return t; }
  Point operator*(Point p) const {
    const Point qv{qx, qy, qz};
    Point uv = cross(qv, p);
    uv = uv + uv;
    const Point r = (p + uv * qw) + cross(qv, uv);
    return r + t;
  }
};

// A.2 grid helpers
static inline GIdx getGridIndexFromPoint(Point p, float inv) {
  return {(int64_t)std::floor(p.x * inv + kEpsilon), (int64_t)std::floor(p.y * inv + kEpsilon),
          (int64_t)std::floor(p.z * inv + kEpsilon)};
}
static inline GIdx getGridIndexFromScaledPoint(Point s) {
  return {(int64_t)std::floor(s.x + kEpsilon), (int64_t)std::floor(s.y + kEpsilon),
          (int64_t)std::floor(s.z + kEpsilon)};
}
static inline Point getCenterPointFromGridIndex(GIdx i, float size) {
  return {((float)i.x + 0.5f) * size, ((float)i.y + 0.5f) * size, ((float)i.z + 0.5f) * size};
}
static inline Point getOriginPointFromGridIndex(BIdx i, float size) {
  return {(float)i.x * size, (float)i.y * size, (float)i.z * size};
}
static inline BIdx getBlockIndexFromGlobalVoxelIndex(GIdx g, float vps_inv) {
  return {(int32_t)std::floor((float)g.x * vps_inv), (int32_t)std::floor((float)g.y * vps_inv),
          (int32_t)std::floor((float)g.z * vps_inv)};
}
static inline void getLocalFromGlobalVoxelIndex(GIdx g, int vps, int* lx, int* ly, int* lz) {
  const int64_t offset = int64_t(1) << 31;  // voxblox: positive offset then bit-and (vps power of two)
  *lx = (int)((g.x + offset) & (vps - 1));
  *ly = (int)((g.y + offset) & (vps - 1));
  *lz = (int)((g.z + offset) & (vps - 1));
}
// AnyIndexHash / LongIndexHash: 64-bit modular arithmetic truncated to 32 bits (A.2)
static inline size_t indexHash(int64_t x, int64_t y, int64_t z) {
  constexpr size_t sl = 17191;
  constexpr size_t sl2 = sl * sl;
  return static_cast<unsigned int>((size_t)x + (size_t)y * sl + (size_t)z * sl2);
}
struct LongIndexHash { size_t operator()(const GIdx& i) const { return indexHash(i.x, i.y, i.z); } };
struct AnyIndexHash { size_t operator()(const BIdx& i) const { return indexHash(i.x, i.y, i.z); } };

// A.3 ThreadSafeIndex
struct ThreadSafeIndex {
  std::atomic<size_t> atomic_idx{0};
  size_t n = 0;
  int mode = KSG_ORDER_MIXED;
  size_t groups = 0;
  std::vector<size_t> sorted;
  ThreadSafeIndex(int mode_, const std::vector<Point>& pts) : n(pts.size()), mode(mode_) {
    groups = n / 1024;  // step_size_ = 1 << 10
    if (mode == KSG_ORDER_SORTED) {
      // voxblox SortedThreadSafeIndex: sort (index, squaredNorm) by squaredNorm. std::sort is not
      // stable upstream; the oracle fixes ties by ascending index (canonical).
      std::vector<std::pair<size_t, float>> v(n);
      for (size_t i = 0; i < n; ++i) v[i] = {i, squaredNorm(pts[i])};
      std::stable_sort(v.begin(), v.end(),
                       [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) { return a.second < b.second; });
      sorted.resize(n);
      for (size_t i = 0; i < n; ++i) sorted[i] = v[i].first;
    }
  }
  size_t impl(size_t seq) const {
    if (mode == KSG_ORDER_SORTED) return sorted[seq];
    if (groups * 1024 <= seq) return seq;
    return (seq % groups) * 1024 + seq / groups;
  }
  bool getNextIndex(size_t* idx) {
    const size_t seq = atomic_idx.fetch_add(1);
    if (seq >= n) return false;
    *idx = impl(seq);
    return true;
  }
};

// A.4 ApproxHashSet<20, 10000>
struct ApproxHashSet {
  static constexpr size_t kBits = 20;
  static constexpr size_t kSize = size_t(1) << kBits;
  static constexpr size_t kMask = kSize - 1;
  static constexpr size_t kFullReset = 10000;
  size_t offset = 0;
  std::vector<std::atomic<size_t>> table;
  ApproxHashSet() : table(kSize) { clear(); }
  void clear() {
    for (auto& v : table) v.store(0, std::memory_order_relaxed);
    offset = 0;
    table[offset].store(std::numeric_limits<size_t>::max(), std::memory_order_relaxed);
  }
  bool replaceHash(size_t hash) {
    const size_t k = (hash + offset) & kMask;
    if (table[k].load(std::memory_order_relaxed) == hash + offset) return false;
    table[k].store(hash + offset, std::memory_order_relaxed);
    return true;
  }
  bool replaceHash(const GIdx& i) { return replaceHash(indexHash(i.x, i.y, i.z)); }
  void resetApproxSet() {
    if (++offset >= kFullReset) clear();
  }
};

// A.7 RayCaster
static inline int signum(float x) { return (0.0f < x) - (x < 0.0f); }
struct RayCaster {
  GIdx curr{0, 0, 0};
  int sign[3] = {0, 0, 0};
  float t_next[3] = {0, 0, 0};
  float t_step[3] = {0, 0, 0};
  uint64_t step = 0;
  uint64_t length_in_steps = 0;

  RayCaster(Point origin, Point point_G, bool is_clearing, bool carving, float max_len, float vsi,
            float trunc, bool cast_from_origin = true) {
    const Point unit_ray = normalized(point_G - origin);
    Point ray_start, ray_end;
    if (is_clearing) {
      float ray_length = norm(point_G - origin);
      ray_length = std::min(std::max(ray_length - trunc, 0.0f), max_len);
      ray_end = origin + unit_ray * ray_length;
      ray_start = carving ? origin : ray_end;
    } else {
      ray_end = point_G + unit_ray * trunc;
      ray_start = carving ? origin : (point_G - unit_ray * trunc);
    }
    const Point start_scaled = ray_start * vsi;
    const Point end_scaled = ray_end * vsi;
    if (cast_from_origin) setup(start_scaled, end_scaled); else setup(end_scaled, start_scaled);
  }
  void setup(Point s, Point e) {
    if (std::isnan(s.x) || std::isnan(s.y) || std::isnan(s.z) || std::isnan(e.x) || std::isnan(e.y) || std::isnan(e.z)) {
      length_in_steps = 0;
      step = 0;
      return;
    }
    curr = getGridIndexFromScaledPoint(s);
    const GIdx endi = getGridIndexFromScaledPoint(e);
    step = 0;
    length_in_steps = (uint64_t)(std::llabs(endi.x - curr.x) + std::llabs(endi.y - curr.y) + std::llabs(endi.z - curr.z));
    const Point r = e - s;
    sign[0] = signum(r.x); sign[1] = signum(r.y); sign[2] = signum(r.z);
    const float corr[3] = {(float)std::max(0, sign[0]), (float)std::max(0, sign[1]), (float)std::max(0, sign[2])};
    const float shifted[3] = {s.x - (float)curr.x, s.y - (float)curr.y, s.z - (float)curr.z};
    const float rr[3] = {r.x, r.y, r.z};
    for (int k = 0; k < 3; ++k) {
      const float dist_to_boundary = corr[k] - shifted[k];
      // upstream guards with (std::abs(r) < 0.0) ? 2.0 : ..., which is never true (A.7)
      t_next[k] = dist_to_boundary / rr[k];
      t_step[k] = (float)sign[k] / rr[k];
    }
  }
  bool nextRayIndex(GIdx* out) {
    if (step++ > length_in_steps) return false;
    *out = curr;
    int k = 0;  // Eigen minCoeff: first minimum
    if (t_next[1] < t_next[k]) k = 1;
    if (t_next[2] < t_next[k]) k = 2;
    int64_t* c = (k == 0) ? &curr.x : (k == 1) ? &curr.y : &curr.z;
    *c += sign[k];
    t_next[k] += t_step[k];
    return true;
  }
};

// voxblox Color::blendTwoColors (A.6)
static inline Color blendTwoColors(Color c1, float w1, Color c2, float w2) {
  const float total = w1 + w2;
  w1 /= total;
  w2 /= total;
  Color o;
  o.r = static_cast<uint8_t>(std::round(c1.r * w1 + c2.r * w2));
  o.g = static_cast<uint8_t>(std::round(c1.g * w1 + c2.g * w2));
  o.b = static_cast<uint8_t>(std::round(c1.b * w1 + c2.b * w2));
  o.a = static_cast<uint8_t>(std::round(c1.a * w1 + c2.a * w2));
  return o;
}

// voxblox rainbowColorMap (voxblox/core/color.h), used by ColorMode::kSemanticProbability base.cpp:181-185
static inline Color rainbowColorMap(double h) {
  Color color;
  color.a = 255;
  double s = 1.0, v = 1.0;
  h -= std::floor(h);
  h *= 6;
  int i = (int)std::floor(h);
  double f = h - i;
  if (!(i & 1)) f = 1 - f;
  const double m = v * (1 - s);
  const double n = v * (1 - s * f);
  switch (i) {
    case 6:
    case 0: color.r = (uint8_t)(255 * v); color.g = (uint8_t)(255 * n); color.b = (uint8_t)(255 * m); break;
    case 1: color.r = (uint8_t)(255 * n); color.g = (uint8_t)(255 * v); color.b = (uint8_t)(255 * m); break;
    case 2: color.r = (uint8_t)(255 * m); color.g = (uint8_t)(255 * v); color.b = (uint8_t)(255 * n); break;
    case 3: color.r = (uint8_t)(255 * m); color.g = (uint8_t)(255 * n); color.b = (uint8_t)(255 * v); break;
    case 4: color.r = (uint8_t)(255 * n); color.g = (uint8_t)(255 * m); color.b = (uint8_t)(255 * v); break;
    case 5: color.r = (uint8_t)(255 * v); color.g = (uint8_t)(255 * m); color.b = (uint8_t)(255 * n); break;
    default: color.r = 255; color.g = 127; color.b = 127; break;
  }
  return color;
}

// ---------------------------------------------------------------------------------------------
// A.1 Layer / Block (subset) with run-time class count
// ---------------------------------------------------------------------------------------------
struct TsdfVoxel { float distance = 0.0f; float weight = 0.0f; Color color; };
struct Block {
  BIdx index;
  Point origin;
  bool updated = false;
  std::vector<TsdfVoxel> tsdf;        // vps^3
  std::vector<uint8_t> sem_label;     // semantic_voxel.h:17  label = 0
  std::vector<float> sem_priors;      // semantic_voxel.h:21-23 priors = -0.60205999132
  std::vector<Color> sem_color;       // semantic_voxel.h:26  Gray
  Block(int vps, int C, BIdx idx, Point org) : index(idx), origin(org) {
    const size_t V = (size_t)vps * vps * vps;
    tsdf.resize(V);
    sem_label.assign(V, 0);
    sem_priors.assign(V * C, (float)-0.60205999132);
    Color gray; gray.r = 127; gray.g = 127; gray.b = 127; gray.a = 255;
    sem_color.assign(V, gray);
  }
};
// The reference keeps two layers (TSDF and semantic) with identical geometry that are always
// allocated together (fast.cpp:124-131, merged.cpp:315-321); the oracle stores both in one Block.
typedef std::unordered_map<BIdx, std::shared_ptr<Block>, AnyIndexHash> BlockHashMap;

struct Mutexes {  // ApproxHashArray<12, std::mutex, GlobalIndex, LongIndexHash>  base.h:64-66
  std::vector<std::mutex> m;
  Mutexes() : m(4096) {}
  std::mutex& get(const GIdx& i) { return m[indexHash(i.x, i.y, i.z) & 4095]; }
};

struct Integrator {
  ksg_config cfg;
  bool canonical_merged = true;
  int C = 21;
  int vps = 16;
  float voxel_size, voxel_size_inv, block_size, vps_inv;
  float log_match, log_non_match;
  std::vector<float> L;  // C x C row-major semantic_log_likelihood_ (base.cpp:93-128)
  BlockHashMap layer;
  BlockHashMap temp_block_map;  // base.h:203-204 (+ the voxblox TSDF twin)
  std::mutex temp_block_mutex;
  Mutexes mutexes;
  ApproxHashSet start_voxel_approx_set, voxel_observed_approx_set;  // fast.h:114-130
  int64_t reset_counter = 0;  // function-static in fast.cpp:165 (per instance here)
  std::unordered_map<uint32_t, uint8_t> color_to_label;  // color.cpp:42-67
  std::vector<BIdx> last_updated;
  std::atomic<int64_t> n_updates{0}, n_rays{0}, n_steps{0}, n_valid{0};
  std::string error;
  double last_integrate_seconds = 0.0;
  bool trace_fast = false;  // debug: record (point index, #updates) of every cast ray
  std::vector<int64_t> trace_point, trace_updates;  // span of the reference timers integrate/fast + inserting_missed_blocks

  explicit Integrator(const ksg_config& c, bool canonical) : cfg(c), canonical_merged(canonical) {
    C = c.num_labels;
    vps = c.voxels_per_side;
    voxel_size = c.voxel_size;
    // voxblox Layer: inverses computed as 1.0 / x (double) then stored as float (A.1);
    // base.cpp:87-89 does the same.
    voxel_size_inv = (float)(1.0 / voxel_size);
    block_size = voxel_size * vps;
    vps_inv = (float)(1.0f / (float)vps);
    // voxblox TsdfIntegratorBase ctor: allow_clear is switched off when carving is disabled (found by running the
    // reference's own sources against the stand-in voxblox, oracle/ref_hybrid.cpp)
    if (cfg.allow_clear && !cfg.voxel_carving_enabled) cfg.allow_clear = 0;
    setSemanticProbabilities();
  }

  // base.cpp:93-128
  void setSemanticProbabilities() {
    const float match = cfg.semantic_measurement_probability;
    const float non_match = 1.0f - cfg.semantic_measurement_probability;
    log_match = std::log(match);
    log_non_match = std::log(non_match);
    L.assign((size_t)C * C, log_non_match);
    for (int i = 0; i < C; ++i) L[(size_t)i * C + i] = log_match;
    for (int i = 0; i < C; ++i) L[(size_t)i * C + 0] = 0.0f;  // col(kUnknownSemanticLabelId).setZero()
  }

  // base.cpp:205-254 and its voxblox twin TsdfIntegratorBase::allocateStorageAndGetVoxelPtr (A.6)
  Block* allocateStorageAndGetBlock(const GIdx& g, std::shared_ptr<Block>* last_block, BIdx* last_idx, size_t* lin) {
    const BIdx bidx = getBlockIndexFromGlobalVoxelIndex(g, vps_inv);
    if (!(*last_block) || !(bidx == *last_idx)) {
      auto it = layer.find(bidx);
      *last_block = (it == layer.end()) ? nullptr : it->second;
      *last_idx = bidx;
    }
    if (!(*last_block)) {
      std::lock_guard<std::mutex> lock(temp_block_mutex);
      auto it = temp_block_map.find(bidx);
      if (it != temp_block_map.end()) {
        *last_block = it->second;
      } else {
        auto ins = temp_block_map.emplace(bidx, std::make_shared<Block>(vps, C, bidx, getOriginPointFromGridIndex(bidx, block_size)));
        *last_block = ins.first->second;
      }
    }
    (*last_block)->updated = true;  // base.cpp:248
    int lx, ly, lz;
    getLocalFromGlobalVoxelIndex(g, vps, &lx, &ly, &lz);
    *lin = (size_t)lx + (size_t)vps * ((size_t)ly + (size_t)vps * (size_t)lz);
    return last_block->get();
  }
  // base.cpp:257-265
  void updateLayerWithStoredBlocks() {
    for (auto& kv : temp_block_map) layer.insert(kv);
    temp_block_map.clear();
  }

  // A.6
  bool isPointValid(Point p_C, bool freespace, bool* is_clearing) const {
    const float d = norm(p_C);
    if (d < cfg.min_ray_length_m) return false;
    if (d > cfg.max_ray_length_m) {
      if (cfg.allow_clear || freespace) { *is_clearing = true; return true; }
      return false;
    }
    *is_clearing = freespace;
    return true;
  }
  float getVoxelWeight(Point p_C) const {
    if (cfg.use_const_weight) return 1.0f;
    const float z = std::abs(p_C.z);
    if (z > kEpsilon) return 1.0f / (z * z);
    return 0.0f;
  }
  bool isSemanticLabelValid(uint8_t l) const { return !cfg.dynamic_label[l]; }  // base.h:170-175

  static float computeDistance(Point origin, Point point_G, Point voxel_center) {
    const Point v_voxel_origin = voxel_center - origin;
    const Point v_point_origin = point_G - origin;
    const float dist_G = norm(v_point_origin);
    const float dist_G_V = dot(v_voxel_origin, v_point_origin) / dist_G;
    return dist_G - dist_G_V;
  }
  // voxblox TsdfIntegratorBase::updateTsdfVoxel (A.6); lock = mutexes_.get(global_voxel_idx)
  void updateTsdfVoxel(Point origin, Point point_G, const GIdx& g, Color color, float weight, TsdfVoxel* v, bool lock) {
    const Point center = getCenterPointFromGridIndex(g, voxel_size);
    const float sdf = computeDistance(origin, point_G, center);
    float updated_weight = weight;
    const float dropoff_epsilon = voxel_size;
    const float trunc = cfg.default_truncation_distance;
    if (cfg.use_weight_dropoff && sdf < -dropoff_epsilon) {
      updated_weight = weight * (trunc + sdf) / (trunc - dropoff_epsilon);
      updated_weight = std::max(updated_weight, 0.0f);
    }
    if (cfg.use_sparsity_compensation_factor) {
      if (std::abs(sdf) < trunc) updated_weight *= cfg.sparsity_compensation_factor;
    }
    std::unique_lock<std::mutex> lk;
    if (lock) lk = std::unique_lock<std::mutex>(mutexes.get(g));
    const float new_weight = v->weight + updated_weight;
    if (new_weight < kEpsilon) return;
    const float new_sdf = (sdf * updated_weight + v->distance * v->weight) / new_weight;
    if (std::abs(sdf) < trunc) v->color = blendTwoColors(v->color, v->weight, color, updated_weight);
    v->distance = (new_sdf > 0.0f) ? std::min(trunc, new_sdf) : std::max(-trunc, new_sdf);
    v->weight = std::min(cfg.max_weight, new_weight);
  }

  // base.cpp:136-194 (updateSemanticVoxel) with base.cpp:283-314, 352-367, 370-380 inlined.
  // `freq` has C entries.  Summation order of the C x C mat-vec is fixed: j ascending, one multiply
  // and one add per term, no FMA (A.9).
  void updateSemanticVoxel(const GIdx& g, const float* freq, Block* b, size_t lin, bool lock) {
    std::unique_lock<std::mutex> lk;
    if (lock) lk = std::unique_lock<std::mutex>(mutexes.get(g));
    float* prior = &b->sem_priors[lin * C];
    for (int i = 0; i < C; ++i) {
      float acc = 0.0f;
      const float* Li = &L[(size_t)i * C];
      for (int j = 0; j < C; ++j) acc = acc + Li[j] * freq[j];
      prior[i] = prior[i] + acc;
    }
    int label = 0;  // maxCoeff: first maximum wins (base.cpp:366)
    float best = prior[0];
    for (int i = 1; i < C; ++i) if (prior[i] > best) { best = prior[i]; label = i; }
    b->sem_label[lin] = (uint8_t)label;
    // base.cpp:370-380 + color.cpp:84-94 (miss -> HashableColor() = (0,0,0,0))
    Color sc;
    if (cfg.label_color_known[label]) { sc.r = cfg.label_color[label][0]; sc.g = cfg.label_color[label][1]; sc.b = cfg.label_color[label][2]; sc.a = cfg.label_color[label][3]; }
    b->sem_color[lin] = sc;
    switch (cfg.color_mode) {  // base.cpp:172-191
      case KSG_COLOR_MODE_COLOR: break;
      case KSG_COLOR_MODE_SEMANTIC: b->tsdf[lin].color = sc; break;
      case KSG_COLOR_MODE_SEMANTIC_PROBABILITY: b->tsdf[lin].color = rainbowColorMap(std::exp(prior[label])); break;
      default: break;
    }
  }

  uint8_t labelFromColor(uint8_t r, uint8_t g, uint8_t b) const {  // color.cpp:69-82, alpha forced to 255
    auto it = color_to_label.find(((uint32_t)r << 16) | ((uint32_t)g << 8) | b);
    return it == color_to_label.end() ? 0 : it->second;
  }

  // ------------------------------------------------------------------------------------------
  // fast.cpp:57-143
  // ------------------------------------------------------------------------------------------
  void integrateSemanticFunction(const Transformation& T_G_C, const std::vector<Point>& points_C,
                                 const std::vector<Color>& colors, const std::vector<uint8_t>& labels,
                                 bool freespace, ThreadSafeIndex* index_getter, bool lock) {
    size_t point_idx;
    std::vector<float> freq(C);
    int64_t updates = 0, rays = 0, valid = 0;
    while (index_getter->getNextIndex(&point_idx)) {  // max_integration_time_s = FLT_MAX: never expires
      const Point point_C = points_C[point_idx];
      const Color color = colors[point_idx];
      const uint8_t semantic_label = labels[point_idx];
      bool is_clearing;
      if (!isPointValid(point_C, freespace, &is_clearing) || !isSemanticLabelValid(semantic_label)) continue;
      ++valid;
      const Point origin = T_G_C.getPosition();
      const Point point_G = T_G_C * point_C;
      GIdx global_voxel_idx = getGridIndexFromPoint(point_G, cfg.start_voxel_subsampling_factor * voxel_size_inv);
      if (!start_voxel_approx_set.replaceHash(global_voxel_idx)) continue;
      ++rays;
      const int64_t updates_before = updates;
      RayCaster ray_caster(origin, point_G, is_clearing, cfg.voxel_carving_enabled, cfg.max_ray_length_m,
                           voxel_size_inv, cfg.default_truncation_distance, /*cast_from_origin=*/false);
      int64_t consecutive_ray_collisions = 0;
      std::shared_ptr<Block> block = nullptr;
      BIdx block_idx{0, 0, 0};
      while (ray_caster.nextRayIndex(&global_voxel_idx)) {
        if (!voxel_observed_approx_set.replaceHash(global_voxel_idx)) ++consecutive_ray_collisions;
        else consecutive_ray_collisions = 0;
        if (consecutive_ray_collisions > cfg.max_consecutive_ray_collisions) break;
        size_t lin;
        Block* b = allocateStorageAndGetBlock(global_voxel_idx, &block, &block_idx, &lin);
        const float weight = getVoxelWeight(point_C);
        updateTsdfVoxel(origin, point_G, global_voxel_idx, color, weight, &b->tsdf[lin], lock);
        std::fill(freq.begin(), freq.end(), 0.0f);
        freq[semantic_label] += 1.0f;  // fast.cpp:132-135
        updateSemanticVoxel(global_voxel_idx, freq.data(), b, lin, lock);
        ++updates;
      }
      if (trace_fast) { trace_point.push_back((int64_t)point_idx); trace_updates.push_back(updates - updates_before); }
    }
    n_updates += updates; n_rays += rays; n_valid += valid;
  }

  // fast.cpp:145-199 (the colour->label loop :152-158 is done by the caller-side wrappers below)
  void integrateFast(const Transformation& T, const std::vector<Point>& pts, const std::vector<Color>& colors,
                     const std::vector<uint8_t>& labels, bool freespace) {
    if ((++reset_counter) >= cfg.clear_checks_every_n_frames) {
      reset_counter = 0;
      start_voxel_approx_set.resetApproxSet();
      voxel_observed_approx_set.resetApproxSet();
    }
    ThreadSafeIndex index_getter(cfg.integration_order_mode, pts);
    trace_point.clear(); trace_updates.clear();
    const int threads = std::max(1, cfg.integrator_threads);
    if (threads == 1) {
      integrateSemanticFunction(T, pts, colors, labels, freespace, &index_getter, false);
    } else {
      std::list<std::thread> ts;
      for (int i = 0; i < threads; ++i)
        ts.emplace_back([&]() { integrateSemanticFunction(T, pts, colors, labels, freespace, &index_getter, true); });
      for (auto& t : ts) t.join();
    }
    collectUpdated();
    updateLayerWithStoredBlocks();
  }

  // ------------------------------------------------------------------------------------------
  // merged.cpp:97-329 (+ voxblox bundleRays A.5)
  // ------------------------------------------------------------------------------------------
  struct Bundle { GIdx key; std::vector<size_t> pts; };

  void integrateVoxel(const Transformation& T_G_C, const std::vector<Point>& points_C,
                      const std::vector<uint8_t>& labels, bool clearing_ray, const GIdx& key,
                      const std::vector<size_t>& pt_indices,
                      const std::unordered_map<GIdx, size_t, LongIndexHash>* voxel_keys, bool lock,
                      std::vector<float>& freq, int64_t* updates) {
    if (pt_indices.empty()) return;
    const Point origin = T_G_C.getPosition();
    Color merged_color;  // HashableColor() = (0,0,0,0); per-point colours are the unfilled hash_colors (merged.cpp:70)
    Point merged_point_C{0.0f, 0.0f, 0.0f};
    float merged_weight = 0.0f;
    std::fill(freq.begin(), freq.end(), 0.0f);
    for (const size_t pt_idx : pt_indices) {
      const Point point_C = points_C[pt_idx];
      const Color color;  // (0,0,0,0)
      const float point_weight = getVoxelWeight(point_C);
      if (point_weight < kEpsilon) continue;
      merged_point_C = (merged_point_C * merged_weight + point_C * point_weight) / (merged_weight + point_weight);
      merged_color = blendTwoColors(merged_color, merged_weight, color, point_weight);
      merged_weight += point_weight;
      freq[labels[pt_idx]] += 1.0f;
      if (clearing_ray) break;  // only take first point when clearing
    }
    const Point merged_point_G = T_G_C * merged_point_C;
    RayCaster ray_caster(origin, merged_point_G, clearing_ray, cfg.voxel_carving_enabled, cfg.max_ray_length_m,
                         voxel_size_inv, cfg.default_truncation_distance);
    GIdx g;
    std::shared_ptr<Block> block = nullptr;
    BIdx block_idx{0, 0, 0};
    while (ray_caster.nextRayIndex(&g)) {
      if (cfg.enable_anti_grazing) {  // merged.cpp:306-313
        if ((clearing_ray || !(g == key)) && voxel_keys->find(g) != voxel_keys->end()) continue;
      }
      size_t lin;
      Block* b = allocateStorageAndGetBlock(g, &block, &block_idx, &lin);
      updateTsdfVoxel(origin, merged_point_G, g, merged_color, merged_weight, &b->tsdf[lin], lock);
      updateSemanticVoxel(g, freq.data(), b, lin, lock);
      ++(*updates);
    }
  }

  void integrateMerged(const Transformation& T, const std::vector<Point>& pts, const std::vector<uint8_t>& labels, bool freespace) {
    ThreadSafeIndex index_getter(cfg.integration_order_mode, pts);
    // bundleRays (A.5), single thread.
    //  canonical mode: bundles iterate in first-insertion order (what the GPU path reproduces);
    //  faithful  mode: a real std::unordered_map<LongIndex, vector, LongIndexHash> is filled in the
    //                  same order and iterated in libstdc++'s order (merged.cpp:210-231).
    std::vector<Bundle> vox_bundles, clr_bundles;
    std::unordered_map<GIdx, size_t, LongIndexHash> vox_index, clr_index;
    std::unordered_map<GIdx, std::vector<size_t>, LongIndexHash> f_vox, f_clr;
    size_t point_idx;
    int64_t valid = 0;
    while (index_getter.getNextIndex(&point_idx)) {
      const Point p = pts[point_idx];
      bool is_clearing;
      if (!isPointValid(p, freespace, &is_clearing)) continue;
      ++valid;
      const Point point_G = T * p;
      const GIdx v = getGridIndexFromPoint(point_G, voxel_size_inv);
      auto& index = is_clearing ? clr_index : vox_index;
      auto& bundles = is_clearing ? clr_bundles : vox_bundles;
      auto it = index.find(v);
      if (it == index.end()) { index.emplace(v, bundles.size()); bundles.push_back({v, {point_idx}}); }
      else bundles[it->second].pts.push_back(point_idx);
      if (!canonical_merged) (is_clearing ? f_clr : f_vox)[v].push_back(point_idx);
    }
    n_valid += valid;
    const int threads = std::max(1, cfg.integrator_threads);
    for (int pass = 0; pass < 2; ++pass) {  // merged.cpp:126-144
      const bool clearing = pass == 1;
      std::vector<const GIdx*> keys;
      std::vector<const std::vector<size_t>*> lists;
      if (canonical_merged) {
        for (auto& b : (clearing ? clr_bundles : vox_bundles)) { keys.push_back(&b.key); lists.push_back(&b.pts); }
      } else {
        for (auto& kv : (clearing ? f_clr : f_vox)) { keys.push_back(&kv.first); lists.push_back(&kv.second); }
      }
      n_rays += (int64_t)keys.size();
      auto worker = [&](int thread_idx, bool lock) {  // merged.cpp:200-232
        std::vector<float> freq(C);
        int64_t updates = 0;
        for (size_t i = 0; i < keys.size(); ++i) {
          if (((i + thread_idx + 1) % threads) == 0)
            integrateVoxel(T, pts, labels, clearing, *keys[i], *lists[i], &vox_index, lock, freq, &updates);
        }
        n_updates += updates;
      };
      if (threads == 1) worker(0, false);
      else {
        std::list<std::thread> ts;
        for (int i = 0; i < threads; ++i) ts.emplace_back(worker, i, true);
        for (auto& t : ts) t.join();
      }
      if (pass == 0) last_updated.clear();
      collectUpdated(/*append=*/true);
      updateLayerWithStoredBlocks();  // merged.cpp:193-196
    }
  }

  void collectUpdated(bool append = false) {
    if (!append) last_updated.clear();
    for (auto& kv : layer) if (kv.second->updated) { last_updated.push_back(kv.first); kv.second->updated = false; }
    for (auto& kv : temp_block_map) if (kv.second->updated) { last_updated.push_back(kv.first); kv.second->updated = false; }
  }

  std::vector<BIdx> sortedBlockIndices() const {
    std::vector<BIdx> v;
    v.reserve(layer.size());
    for (auto& kv : layer) v.push_back(kv.first);
    std::sort(v.begin(), v.end(), [](const BIdx& a, const BIdx& b) {
      if (a.z != b.z) return a.z < b.z;
      if (a.y != b.y) return a.y < b.y;
      return a.x < b.x;
    });
    return v;
  }
};

}  // namespace kso

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------------------------
using namespace kso;

static Transformation makeT(const float* T) { return Transformation{T[0], T[1], T[2], T[3], Point{T[4], T[5], T[6]}}; }

extern "C" {

void* kso_create(const ksg_config* cfg, int canonical_merged) {
  if (!cfg || cfg->num_labels < 2 || cfg->num_labels > 256) return nullptr;
  const int v = cfg->voxels_per_side;
  if (v <= 0 || (v & (v - 1))) return nullptr;
  const float p = cfg->semantic_measurement_probability;
  if (!(p > 0.0f && p < 1.0f) || !(std::log(p) > std::log(1.0f - p))) return nullptr;  // base.cpp:98-107
  return new Integrator(*cfg, canonical_merged != 0);
}
void kso_destroy(void* h) { delete (Integrator*)h; }

int kso_set_color_to_label(void* hh, const uint8_t* rgb, const uint8_t* labels, int n) {
  Integrator* h = (Integrator*)hh;
  h->color_to_label.clear();
  for (int i = 0; i < n; ++i)
    h->color_to_label[((uint32_t)rgb[3 * i] << 16) | ((uint32_t)rgb[3 * i + 1] << 8) | rgb[3 * i + 2]] = labels[i];
  return 0;
}

static void fillStats(Integrator* h, int64_t n, ksg_frame_stats* s, int64_t u0, int64_t r0, int64_t v0) {
  if (!s) return;
  std::memset(s, 0, sizeof(*s));
  s->points_in = n;
  s->points_valid = h->n_valid - v0;
  s->rays_cast = h->n_rays - r0;
  s->voxel_updates = h->n_updates - u0;
  s->blocks_allocated = (int64_t)h->layer.size();
  {
    std::vector<BIdx> v = h->last_updated;  // merged appends per pass: count unique blocks
    std::sort(v.begin(), v.end(), [](const BIdx& a, const BIdx& b) { return a.z != b.z ? a.z < b.z : (a.y != b.y ? a.y < b.y : a.x < b.x); });
    v.erase(std::unique(v.begin(), v.end()), v.end());
    s->blocks_touched = (int64_t)v.size();
  }
}

int kso_integrate_points(void* hh, const float* T_G_C, const float* xyz, const uint8_t* rgba,
                         const uint8_t* labels_in, int64_t n, int freespace, ksg_frame_stats* stats) {
  Integrator* h = (Integrator*)hh;
  const Transformation T = makeT(T_G_C);
  std::vector<Point> pts((size_t)n);
  std::vector<Color> colors((size_t)n);
  std::vector<uint8_t> labels((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    pts[i] = Point{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    if (labels_in) labels[i] = labels_in[i];
    else if (rgba) labels[i] = h->labelFromColor(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2]);  // fast.cpp:152-158
    else labels[i] = 0;
    if (rgba) { colors[i].r = rgba[4 * i]; colors[i].g = rgba[4 * i + 1]; colors[i].b = rgba[4 * i + 2]; colors[i].a = rgba[4 * i + 3]; }
    else {  // depth+label entry: the point colour is the label's colour (the semantic image pixel)
      const uint8_t l = labels[i];
      if (h->cfg.label_color_known[l]) {
        colors[i].r = h->cfg.label_color[l][0]; colors[i].g = h->cfg.label_color[l][1];
        colors[i].b = h->cfg.label_color[l][2]; colors[i].a = h->cfg.label_color[l][3];
      }
    }
    if (labels[i] >= h->C) return KSG_ERR_INVALID_ARGUMENT;  // CHECK_LT fast.cpp:134
  }
  const int64_t u0 = h->n_updates, r0 = h->n_rays, v0 = h->n_valid;
  const auto t0 = std::chrono::steady_clock::now();
  if (h->cfg.integrator_type == KSG_INTEGRATOR_FAST) h->integrateFast(T, pts, colors, labels, freespace != 0);
  else h->integrateMerged(T, pts, labels, freespace != 0);
  h->last_integrate_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fillStats(h, n, stats, u0, r0, v0);
  return 0;
}

// depth_map_to_pointcloud.h:222-266 (convert<float>) followed by voxblox_ros convertPointcloud's
// finite-point filter, then integratePointCloud.
int64_t kso_backproject_k64(const float* depth, int width, int height, const double* K, float* xyz_out, int32_t* pix_out) {
  const float center_x = (float)K[2], center_y = (float)K[3];   // float center_x = model_.cx() (double)
  const double unit_scaling = 1.0;  // DepthTraits<float>::toMeters(1)
  const float constant_x = (float)(unit_scaling / K[0]);        // float constant_x = unit_scaling / model_.fx() (double)
  const float constant_y = (float)(unit_scaling / K[1]);
  int64_t n = 0;
  for (int v = 0; v < height; ++v)
    for (int u = 0; u < width; ++u) {
      const float d = depth[(size_t)v * width + u];
      if (!std::isfinite(d)) continue;  // DepthTraits<float>::valid -> NaN point -> dropped
      xyz_out[3 * n + 0] = (u - center_x) * d * constant_x;
      xyz_out[3 * n + 1] = (v - center_y) * d * constant_y;
      xyz_out[3 * n + 2] = d;
      if (pix_out) pix_out[n] = v * width + u;
      ++n;
    }
  return n;
}
int64_t kso_backproject(const float* depth, int width, int height, const float* K, float* xyz_out, int32_t* pix_out) {
  const double K64[4] = {K[0], K[1], K[2], K[3]};
  return kso_backproject_k64(depth, width, height, K64, xyz_out, pix_out);
}

int kso_integrate_depth_k64(void* hh, const float* T_G_C, const float* depth, const uint8_t* label, int width,
                            int height, const double* K, ksg_frame_stats* stats) {
  const size_t P = (size_t)width * height;
  std::vector<float> xyz(3 * P);
  std::vector<int32_t> pix(P);
  const int64_t n = kso_backproject_k64(depth, width, height, K, xyz.data(), pix.data());
  std::vector<uint8_t> labels((size_t)n);
  for (int64_t i = 0; i < n; ++i) labels[i] = label[pix[i]];
  return kso_integrate_points(hh, T_G_C, xyz.data(), nullptr, labels.data(), n, 0, stats);
}
int kso_integrate_depth(void* hh, const float* T_G_C, const float* depth, const uint8_t* label, int width,
                        int height, const float* K, ksg_frame_stats* stats) {
  const double K64[4] = {K[0], K[1], K[2], K[3]};
  return kso_integrate_depth_k64(hh, T_G_C, depth, label, width, height, K64, stats);
}

int64_t kso_num_blocks(void* hh) { return (int64_t)((Integrator*)hh)->layer.size(); }

// Layer::removeAllBlocks() on both layers while the integrator object lives on: the fast integrator's two approximate sets keep their
// contents and offsets (fast.h:114-130 - they are members of the integrator, not of the layers).  Twin of ksg_clear_map.
void kso_clear_map(void* hh) {
  Integrator* in = (Integrator*)hh;
  in->layer.clear();
  in->temp_block_map.clear();
  in->last_updated.clear();
}

int kso_export_blocks(void* hh, int64_t capacity, int32_t* block_index, float* tsdf_distance, float* tsdf_weight,
                      uint8_t* tsdf_rgba, uint8_t* sem_label, float* sem_priors, uint8_t* sem_rgba) {
  Integrator* h = (Integrator*)hh;
  const std::vector<BIdx> idx = h->sortedBlockIndices();
  if ((int64_t)idx.size() > capacity) return KSG_ERR_INVALID_ARGUMENT;
  const size_t V = (size_t)h->vps * h->vps * h->vps;
  const int C = h->C;
  for (size_t b = 0; b < idx.size(); ++b) {
    const Block& blk = *h->layer.at(idx[b]);
    if (block_index) { block_index[3 * b] = idx[b].x; block_index[3 * b + 1] = idx[b].y; block_index[3 * b + 2] = idx[b].z; }
    for (size_t v = 0; v < V; ++v) {
      if (tsdf_distance) tsdf_distance[b * V + v] = blk.tsdf[v].distance;
      if (tsdf_weight) tsdf_weight[b * V + v] = blk.tsdf[v].weight;
      if (tsdf_rgba) { uint8_t* o = &tsdf_rgba[(b * V + v) * 4]; o[0] = blk.tsdf[v].color.r; o[1] = blk.tsdf[v].color.g; o[2] = blk.tsdf[v].color.b; o[3] = blk.tsdf[v].color.a; }
      if (sem_label) sem_label[b * V + v] = blk.sem_label[v];
      if (sem_rgba) { uint8_t* o = &sem_rgba[(b * V + v) * 4]; o[0] = blk.sem_color[v].r; o[1] = blk.sem_color[v].g; o[2] = blk.sem_color[v].b; o[3] = blk.sem_color[v].a; }
    }
    if (sem_priors) std::memcpy(&sem_priors[b * V * C], blk.sem_priors.data(), V * C * sizeof(float));
  }
  return 0;
}

int64_t kso_last_updated_blocks(void* hh, int64_t capacity, int32_t* block_index) {
  Integrator* h = (Integrator*)hh;
  std::vector<BIdx> v = h->last_updated;
  std::sort(v.begin(), v.end(), [](const BIdx& a, const BIdx& b) {
    if (a.z != b.z) return a.z < b.z;
    if (a.y != b.y) return a.y < b.y;
    return a.x < b.x;
  });
  v.erase(std::unique(v.begin(), v.end()), v.end());
  if (block_index && (int64_t)v.size() <= capacity)
    for (size_t i = 0; i < v.size(); ++i) { block_index[3 * i] = v[i].x; block_index[3 * i + 1] = v[i].y; block_index[3 * i + 2] = v[i].z; }
  return (int64_t)v.size();
}

// ---- primitives exposed for the known-answer tests -------------------------------------------
uint64_t kso_index_hash(int64_t x, int64_t y, int64_t z) { return (uint64_t)indexHash(x, y, z); }
uint64_t kso_mixed_index(uint64_t n, uint64_t seq) {
  std::vector<Point> dummy;
  const size_t groups = n / 1024;
  if (groups * 1024 <= seq) return seq;
  return (seq % groups) * 1024 + seq / groups;
}
void kso_transform(const float* T, const float* p, float* out) {
  const Point r = makeT(T) * Point{p[0], p[1], p[2]};
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void kso_grid_index(const float* p, float inv, int64_t* out) {
  const GIdx g = getGridIndexFromPoint(Point{p[0], p[1], p[2]}, inv);
  out[0] = g.x; out[1] = g.y; out[2] = g.z;
}
void kso_block_and_local(const int64_t* g, int vps, int32_t* block, int32_t* local) {
  const GIdx gi{g[0], g[1], g[2]};
  const BIdx b = getBlockIndexFromGlobalVoxelIndex(gi, 1.0f / (float)vps);
  block[0] = b.x; block[1] = b.y; block[2] = b.z;
  getLocalFromGlobalVoxelIndex(gi, vps, &local[0], &local[1], &local[2]);
}
int64_t kso_raycast(const float* origin, const float* point_G, int is_clearing, int carving, float max_len,
                    float vsi, float trunc, int cast_from_origin, int64_t* out, int64_t capacity) {
  RayCaster rc(Point{origin[0], origin[1], origin[2]}, Point{point_G[0], point_G[1], point_G[2]}, is_clearing != 0,
               carving != 0, max_len, vsi, trunc, cast_from_origin != 0);
  GIdx g;
  int64_t n = 0;
  while (rc.nextRayIndex(&g)) {
    if (n < capacity) { out[3 * n] = g.x; out[3 * n + 1] = g.y; out[3 * n + 2] = g.z; }
    ++n;
  }
  return n;
}
// Applies a sequence of n TSDF updates (sdf given through point/centre geometry) to one voxel and
// returns distance / weight / colour: pins the clamp-order behaviour.
void kso_tsdf_update_sequence(const ksg_config* cfg, const float* origin, const float* points_G, const float* weights,
                              const uint8_t* rgba, int n, const int64_t* gidx, float* out_dw, uint8_t* out_rgba) {
  Integrator h(*cfg, true);
  TsdfVoxel v;
  for (int i = 0; i < n; ++i) {
    Color c; c.r = rgba[4 * i]; c.g = rgba[4 * i + 1]; c.b = rgba[4 * i + 2]; c.a = rgba[4 * i + 3];
    h.updateTsdfVoxel(Point{origin[0], origin[1], origin[2]}, Point{points_G[3 * i], points_G[3 * i + 1], points_G[3 * i + 2]},
                      GIdx{gidx[0], gidx[1], gidx[2]}, c, weights[i], &v, false);
  }
  out_dw[0] = v.distance; out_dw[1] = v.weight;
  out_rgba[0] = v.color.r; out_rgba[1] = v.color.g; out_rgba[2] = v.color.b; out_rgba[3] = v.color.a;
}
void kso_log_likelihood(const ksg_config* cfg, float* out_CxC, float* out_lm_ln) {
  Integrator h(*cfg, true);
  std::memcpy(out_CxC, h.L.data(), h.L.size() * sizeof(float));
  out_lm_ln[0] = h.log_match; out_lm_ln[1] = h.log_non_match;
}
// n_obs semantic updates with frequency vectors freq[n_obs][C] applied to a fresh voxel.
void kso_semantic_update_sequence(const ksg_config* cfg, const float* freq, int n_obs, float* out_priors, uint8_t* out_label,
                                  uint8_t* out_sem_rgba, uint8_t* out_tsdf_rgba) {
  Integrator h(*cfg, true);
  ksg_config c1 = *cfg;
  Block b(1, cfg->num_labels, BIdx{0, 0, 0}, Point{0, 0, 0});
  for (int i = 0; i < n_obs; ++i) h.updateSemanticVoxel(GIdx{0, 0, 0}, &freq[(size_t)i * cfg->num_labels], &b, 0, false);
  std::memcpy(out_priors, b.sem_priors.data(), sizeof(float) * cfg->num_labels);
  *out_label = b.sem_label[0];
  out_sem_rgba[0] = b.sem_color[0].r; out_sem_rgba[1] = b.sem_color[0].g; out_sem_rgba[2] = b.sem_color[0].b; out_sem_rgba[3] = b.sem_color[0].a;
  out_tsdf_rgba[0] = b.tsdf[0].color.r; out_tsdf_rgba[1] = b.tsdf[0].color.g; out_tsdf_rgba[2] = b.tsdf[0].color.b; out_tsdf_rgba[3] = b.tsdf[0].color.a;
  (void)c1;
}
// ApproxHashSet scripted test: ops[i] = hash to replaceHash, or UINT64_MAX = resetApproxSet.
void kso_approx_set_script(const uint64_t* ops, int n, uint8_t* results) {
  ApproxHashSet s;
  for (int i = 0; i < n; ++i) {
    if (ops[i] == std::numeric_limits<uint64_t>::max()) { s.resetApproxSet(); results[i] = 2; }
    else results[i] = s.replaceHash((size_t)ops[i]) ? 1 : 0;
  }
}
void kso_blend(const uint8_t* c1, float w1, const uint8_t* c2, float w2, uint8_t* out) {
  Color a; a.r = c1[0]; a.g = c1[1]; a.b = c1[2]; a.a = c1[3];
  Color b; b.r = c2[0]; b.g = c2[1]; b.b = c2[2]; b.a = c2[3];
  const Color o = blendTwoColors(a, w1, b, w2);
  out[0] = o.r; out[1] = o.g; out[2] = o.b; out[3] = o.a;
}
void kso_rainbow(double h, uint8_t* out) { const Color c = rainbowColorMap(h); out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = c.a; }

// libstdc++ probe for tests/test_unordered_map_order.py: insert n distinct voxel keys (int64 xyz triples, first-insertion order)
// into the map type merged.cpp:110-113 uses and report (a) the iteration order as indices into the input and (b) the bucket count
// after every insertion.  This is what `merged` in faithful mode iterates (merged.cpp:210-231).
int64_t kso_unordered_map_order(const int64_t* keys, int64_t n, int64_t* order_out, int64_t* bucket_count_after_insert) {
  std::unordered_map<GIdx, size_t, LongIndexHash> m;
  for (int64_t i = 0; i < n; ++i) {
    m.emplace(GIdx{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]}, (size_t)i);
    if (bucket_count_after_insert) bucket_count_after_insert[i] = (int64_t)m.bucket_count();
  }
  int64_t k = 0;
  for (const auto& kv : m) order_out[k++] = (int64_t)kv.second;
  return k;
}

// debug hooks used by tests/ to validate the parallel observed-set solver against the sequential sets
void kso_trace_fast(void* hh, int enable) { ((Integrator*)hh)->trace_fast = enable != 0; }
int64_t kso_get_fast_trace(void* hh, int64_t capacity, int64_t* point_idx, int64_t* updates) {
  Integrator* h = (Integrator*)hh;
  const int64_t n = (int64_t)h->trace_point.size();
  if (point_idx && updates && n <= capacity) {
    std::memcpy(point_idx, h->trace_point.data(), n * sizeof(int64_t));
    std::memcpy(updates, h->trace_updates.data(), n * sizeof(int64_t));
  }
  return n;
}
// which: 0 = start_voxel_approx_set_, 1 = voxel_observed_approx_set_. out has 2^20 entries.
uint64_t kso_get_approx_set(void* hh, int which, uint64_t* out) {
  Integrator* h = (Integrator*)hh;
  ApproxHashSet& s = which ? h->voxel_observed_approx_set : h->start_voxel_approx_set;
  if (out) for (size_t i = 0; i < ApproxHashSet::kSize; ++i) out[i] = (uint64_t)s.table[i].load();
  return (uint64_t)s.offset;
}

// wall-clock of the last integrate call (timed span = fast.cpp:160-198 / merged.cpp:106-148: the
// colour->label loop and input marshalling are outside), for bench.py's CPU legs
double kso_last_integrate_seconds(void* hh) { return ((Integrator*)hh)->last_integrate_seconds; }

}  // extern "C"
