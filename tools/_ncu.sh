set -x
mkdir -p gpurun_out
bash tools/gpu_r2.sh ncu r3 cfg2 cfg2_j0_cheb_o6_f32
bash tools/gpu_r2.sh ncu r3 cfg3 cfg3_j0_cheb_o13_f64
bash tools/gpu_r2.sh ncu r3 cfg2sorted cfg2_j0_cheb_o6_f32 --order sorted
bash tools/gpu_r2.sh launches r3
ls -la gpurun_out | tail -12
