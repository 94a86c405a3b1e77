// Host-side voxel of the semantic layer as seen by callers of the drop-in integrators.
// Field names, order and default values follow the reference voxel (reference semantic_voxel.h:14-27) because
// downstream code reads them by name; on the device the same information is stored as planes / rows of a tile
//  and copied into this struct by GpuIntegratorCore::copyBlocks.
#pragma once
#include "kimera_semantics/color.h"
#include "kimera_semantics/common.h"

namespace kimera {

// log10(1/4): the value the reference hard-codes as the initial (unnormalised) log-probability of every class
static constexpr SemanticProbability kInitialLogProbability = static_cast<SemanticProbability>(-0.60205999132);

struct SemanticVoxel {
  SemanticLabel semantic_label;          // arg-max class, kUnknownSemanticLabelId until the first observation
  SemanticProbabilities semantic_priors; // one accumulated log-probability per class
  HashableColor color;                   // colour of `semantic_label` in the label -> colour table

  SemanticVoxel()
      : semantic_label(kUnknownSemanticLabelId),
        semantic_priors(SemanticProbabilities::Constant(kInitialLogProbability)),
        color(vxb::Color::Gray()) {}
};

}  // namespace kimera
