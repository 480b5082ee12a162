// leg_f32 kernels, orders 10..16 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(leg_f32_hi, ai::kLeg, float, ai::kSplitOrder + 1, ai::kMaxOrder)
