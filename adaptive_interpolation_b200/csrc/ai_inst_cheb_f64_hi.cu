// cheb_f64 kernels, orders 10..16 (see ai_inst.cuh)
#include "ai_inst.cuh"
AI_DEFINE_UNIT(cheb_f64_hi, ai::kCheb, double, ai::kSplitOrder + 1, ai::kMaxOrder)
