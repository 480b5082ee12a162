// cheb_f32 kernels, orders 0..9 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(cheb_f32_lo, ai::kCheb, float, 0, ai::kSplitOrder)
