// leg_f64 kernels, orders 10..16 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(leg_f64_hi, ai::kLeg, double, ai::kSplitOrder + 1, ai::kMaxOrder)
