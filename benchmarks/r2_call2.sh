#!/bin/bash
# round-2 GPU call 2: fp32/TF32 path (tcgen05 kind::tf32 GEMM, TF32 stem, fused loss) -- tests + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv1x1_gpu.py tests/test_fused_loss_gpu.py tests/test_fused_bn_gpu.py tests/test_flagship_gpu.py -x -q > gpurun_out/r2c2_tests.log 2>&1
tail -15 gpurun_out/r2c2_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
tail -3 gpurun_out/r2c2_bench.err; cat gpurun_out/r2c2_bench.json
timeout 300 python benchmarks/conv1x1_bench.py --dtype fp32 --iters 9 --out gpurun_out/r2c2_conv1x1_fp32.json > gpurun_out/r2c2_conv1x1_fp32.log 2>&1
tail -30 gpurun_out/r2c2_conv1x1_fp32.log
SGP_B200_C1_F32_RESIDENT=0 timeout 300 python benchmarks/conv1x1_bench.py --dtype fp32 --iters 9 > gpurun_out/r2c2_conv1x1_fp32_wide.log 2>&1
tail -30 gpurun_out/r2c2_conv1x1_fp32_wide.log
timeout 200 python benchmarks/profile_bn.py --fp32 > gpurun_out/r2c2_bn_fp32.log 2>&1
cat gpurun_out/r2c2_bn_fp32.log
timeout 200 python benchmarks/stem_conv_bench.py --fp32 > gpurun_out/r2c2_stem_fp32.log 2>&1
cat gpurun_out/r2c2_stem_fp32.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c2_launches_fp32.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --skip-e2e --no-secondary > gpurun_out/r2c2_ncu.log 2>&1
