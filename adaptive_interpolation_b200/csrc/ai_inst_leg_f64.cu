// leg_f64 kernels, orders 0..16 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(leg_f64, ai::kLeg, double)
