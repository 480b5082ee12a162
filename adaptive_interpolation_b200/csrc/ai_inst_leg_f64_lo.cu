// leg_f64 kernels, orders 0..9 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(leg_f64_lo, ai::kLeg, double, 0, ai::kSplitOrder)
