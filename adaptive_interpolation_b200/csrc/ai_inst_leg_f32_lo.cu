// leg_f32 kernels, orders 0..9 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(leg_f32_lo, ai::kLeg, float, 0, ai::kSplitOrder)
