set -x
mkdir -p gpurun_out
bash tools/gpu_r2.sh ncu r3 cfg4f32 cfg4_f1_leg_o7_f32 --log2-points 26
bash tools/gpu_r2.sh ncu r3 cfg4f64 cfg4_f1_leg_o7_f64 --log2-points 26
