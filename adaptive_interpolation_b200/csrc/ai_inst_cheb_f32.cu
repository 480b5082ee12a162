// cheb_f32 kernels, orders 0..16 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(cheb_f32, ai::kCheb, float)
