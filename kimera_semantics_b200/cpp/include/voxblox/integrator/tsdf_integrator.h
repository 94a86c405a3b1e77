// voxblox::TsdfIntegratorBase subset: Config (defaults SURVEY.md A.6) + the pure-virtual integratePointCloud that the
// kimera_semantics integrators override (fast.h:82-86, merged.h:70-73).  MergedTsdfIntegrator is the (empty) base the
// reference's MergedSemanticTsdfIntegrator derives from (merged.h:56-59).
#pragma once
#include <limits>
#include <string>
#include <thread>
#include "voxblox/core/layer.h"
#include "voxblox/core/voxel.h"
namespace voxblox {
class TsdfIntegratorBase {
 public:
  typedef std::shared_ptr<TsdfIntegratorBase> Ptr;
  struct Config {
    float default_truncation_distance = 0.1f;
    float max_weight = 10000.0f;
    bool voxel_carving_enabled = true;
    FloatingPoint min_ray_length_m = 0.1f;
    FloatingPoint max_ray_length_m = 5.0f;
    bool use_const_weight = false;
    bool allow_clear = true;
    bool use_weight_dropoff = true;
    bool use_sparsity_compensation_factor = false;
    float sparsity_compensation_factor = 1.0f;
    size_t integrator_threads = std::thread::hardware_concurrency();
    std::string integration_order_mode = "mixed";
    bool enable_anti_grazing = false;                 // merged
    float start_voxel_subsampling_factor = 2.0f;      // fast
    int max_consecutive_ray_collisions = 2;           // fast
    int clear_checks_every_n_frames = 1;              // fast
    float max_integration_time_s = std::numeric_limits<float>::max();  // fast; ignored by the GPU path
  };
  TsdfIntegratorBase(const Config& config, Layer<TsdfVoxel>* layer) : config_(config), layer_(layer) {
    KSG_CHECK(layer != nullptr);
    voxel_size_ = layer_->voxel_size();
    block_size_ = layer_->block_size();
    voxels_per_side_ = layer_->voxels_per_side();
    voxel_size_inv_ = 1.0 / voxel_size_;
    block_size_inv_ = 1.0 / block_size_;
    voxels_per_side_inv_ = 1.0 / voxels_per_side_;
  }
  virtual ~TsdfIntegratorBase() = default;
  virtual void integratePointCloud(const Transformation& T_G_C, const Pointcloud& points_C, const Colors& colors,
                                   const bool freespace_points = false) = 0;
  const Config& getConfig() const { return config_; }
 protected:
  Config config_;
  Layer<TsdfVoxel>* layer_;
  FloatingPoint voxel_size_, block_size_, voxel_size_inv_, voxels_per_side_inv_, block_size_inv_;
  size_t voxels_per_side_;
};
class MergedTsdfIntegrator : public TsdfIntegratorBase {
 public:
  MergedTsdfIntegrator(const Config& config, Layer<TsdfVoxel>* layer) : TsdfIntegratorBase(config, layer) {}
};
}  // namespace voxblox
