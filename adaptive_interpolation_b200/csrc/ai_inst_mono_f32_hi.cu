// mono_f32 kernels, orders 10..16 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(mono_f32_hi, ai::kMono, float, ai::kSplitOrder + 1, ai::kMaxOrder)
