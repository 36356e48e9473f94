#!/bin/bash
# 2-GPU validation of the copy-engine OSGP gather + reference AD-PSGD arm + bf16 NVLS test
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_gdp_loopback_gpu.py -q -x > gpurun_out/r2c5c_tests_1gpu.log 2>&1
tail -5 gpurun_out/r2c5c_tests_1gpu.log | cut -c1-200
timeout 500 python -m pytest tests/test_multigpu.py -q -k "nvls or kernels_match_simulation or graphed_trainer or schedule_swap or bounded" > gpurun_out/r2c5c_tests_multigpu_n2.log 2>&1
tail -8 gpurun_out/r2c5c_tests_multigpu_n2.log | cut -c1-200
timeout 200 python bench.py --gpus 2 --algo osgp --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c5c_bench_n2_osgp_dma.json 2> gpurun_out/r2c5c_bench_n2_osgp_dma.err
tail -2 gpurun_out/r2c5c_bench_n2_osgp_dma.err | cut -c1-200; cut -c1-200 gpurun_out/r2c5c_bench_n2_osgp_dma.json
SGP_B200_GATHER_DMA=0 timeout 200 python bench.py --gpus 2 --algo osgp --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c5c_bench_n2_osgp_kernel.json 2> gpurun_out/r2c5c_bench_n2_osgp_kernel.err
cut -c1-200 gpurun_out/r2c5c_bench_n2_osgp_kernel.json
timeout 200 python bench.py --gpus 2 --algo sgp --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c5c_bench_n2_sgp.json 2> gpurun_out/r2c5c_bench_n2_sgp.err
cut -c1-200 gpurun_out/r2c5c_bench_n2_sgp.json
timeout 200 python bench.py --gpus 2 --algo osgp --batch-size 32 --steps 40 --warmup 5 --no-secondary > gpurun_out/r2c5c_bench_n2_osgp_bs32.json 2> gpurun_out/r2c5c_bench_n2_osgp_bs32.err
cut -c1-200 gpurun_out/r2c5c_bench_n2_osgp_bs32.json
timeout 200 python bench.py --gpus 2 --algo sgp --batch-size 32 --steps 40 --warmup 5 --no-secondary > gpurun_out/r2c5c_bench_n2_sgp_bs32.json 2> gpurun_out/r2c5c_bench_n2_sgp_bs32.err
cut -c1-200 gpurun_out/r2c5c_bench_n2_sgp_bs32.json
timeout 240 python bench.py --gpus 2 --impl reference --algo adpsgd --steps 20 --warmup 10 > gpurun_out/r2c5c_ref_n2_adpsgd.json 2> gpurun_out/r2c5c_ref_n2_adpsgd.err
grep -v "^$" gpurun_out/r2c5c_ref_n2_adpsgd.err | grep -iv "omp_num\|^\*\*\*\|warn" | tail -8 | cut -c1-250; cat gpurun_out/r2c5c_ref_n2_adpsgd.json | cut -c1-400
