// cheb_f32 kernels, orders 10..16 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(cheb_f32_hi, ai::kCheb, float, ai::kSplitOrder + 1, ai::kMaxOrder)
