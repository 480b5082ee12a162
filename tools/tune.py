#!/usr/bin/env python
"""Build tuning variants of libai_b200.so (compile-time knobs of csrc/ai_kernels.cuh) here, then
sweep them on the GPU box:

    python tools/tune.py build                 # in the container: writes _lib/variant_*.so
    python tools/tune.py run [--tables ...]    # on the GPU: runs tools/sweep.py per variant
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "sel_0_0": ["AI_SEL_LDS_F32=0", "AI_SEL_LDS_F64=0"],
    "sel_4_4": ["AI_SEL_LDS_F32=4", "AI_SEL_LDS_F64=4"],
    "sel_12_10": ["AI_SEL_LDS_F32=12", "AI_SEL_LDS_F64=10"],
    "sel_16_16": ["AI_SEL_LDS_F32=16", "AI_SEL_LDS_F64=16"],
}
SEL_TABLES = "approx__32o13-p1.19209289551e-06,approx__64o13-p1.19209289551e-06,mx__cheb_o13_f32"

VARIANTS = {
    "bin_t128_b4": ["AI_BIN_THREADS=128", "AI_BIN_MIN_BLOCKS=4"],
    "bin_t128_b5_np": ["AI_BIN_THREADS=128", "AI_BIN_MIN_BLOCKS=5", "AI_BIN_PREFETCH=0"],
    "bin_t128_b6_np": ["AI_BIN_THREADS=128", "AI_BIN_MIN_BLOCKS=6", "AI_BIN_PREFETCH=0"],
    "bin_t256_b3_np": ["AI_BIN_THREADS=256", "AI_BIN_MIN_BLOCKS=3", "AI_BIN_PREFETCH=0"],
    "bin_t64_b8": ["AI_BIN_THREADS=64", "AI_BIN_MIN_BLOCKS=8", "AI_BIN_Q=13"],
    "bin_t64_b10_np": ["AI_BIN_THREADS=64", "AI_BIN_MIN_BLOCKS=10", "AI_BIN_PREFETCH=0", "AI_BIN_Q=13"],
}
BIN_TABLES = "gen__cfg3_j0_cheb_o13_f64,mx__cheb_o10_f64,mx__cheb_o12_f64,mx__cheb_o13_f64"

VARIANTS = {
    # tag: defines   (defaults: AI_BLOCK=512 AI_UNROLL=6 AI_UNROLL_F64=5 AI_PREFETCH=1 AI_UNIFORM_PATH=2 AI_PACKED_F32=1)
    "debug_bounds": ["AI_DEBUG_BOUNDS=1"],
    "packed_uniform_only": ["AI_PACKED_F32=2"],
    "packed_off": ["AI_PACKED_F32=0"],
    # CTA shape: the same 16 warps per SM as two or four independent CTAs (round 2)
    "blk256_mb2": ["AI_BLOCK=256", "AI_MIN_BLOCKS=2"],
    "blk128_mb4": ["AI_BLOCK=128", "AI_MIN_BLOCKS=4"],
    # two tiles in flight per thread at the FULL tile size (round 1 only tried it with shrunk tiles)
    "pf2": ["AI_PREFETCH=2"],
}
TABLES = ("approx__32o6-p1.19209289551e-06,approx__32o13-p1.19209289551e-06,gen__cfg3_j0_cheb_o13_f64,"
          "approx__64o7-p2.22044604925e-14,approx__64o6-p1.19209289551e-06,approx__32o3-p1.19209289551e-06")


def main():
    from adaptive_interpolation_b200 import build
    cmd = sys.argv[1] if len(sys.argv) > 1 else "build"
    if cmd == "build":
        only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else list(VARIANTS)
        for tag, defs in VARIANTS.items():
            if tag not in only:
                continue
            print(tag, build.build(defines=defs, out="variant_%s.so" % tag))
        return
    tables = TABLES
    if "--tables" in sys.argv:
        tables = sys.argv[sys.argv.index("--tables") + 1]
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else list(VARIANTS)
    for tag in only:
        lib = os.path.join(ROOT, "adaptive_interpolation_b200", "_lib", "variant_%s.so" % tag)
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, AI_B200_LIB=lib)
        print("=== variant", tag, VARIANTS.get(tag), flush=True)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep.py"), "--tables", tables,
                        "--tag", "tune_" + tag, "--reps", "15"], env=env)


if __name__ == "__main__":
    main()
