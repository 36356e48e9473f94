#!/bin/bash
# quick 2-GPU validation before the 8-GPU call: NVLS rendezvous, loop-back tests, reference AD-PSGD arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gdp_loopback_gpu.py tests/test_flagship_gpu.py tests/test_kernels_gpu.py -q > gpurun_out/r2c5b_tests_1gpu.log 2>&1
tail -6 gpurun_out/r2c5b_tests_1gpu.log | cut -c1-200
timeout 600 python -m pytest tests/test_multigpu.py -q -k "nvls or vmm or hierarchical" > gpurun_out/r2c5b_tests_nvls_n2.log 2>&1
tail -25 gpurun_out/r2c5b_tests_nvls_n2.log | cut -c1-250
timeout 200 python bench.py --gpus 2 --algo ar --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c5b_bench_n2_ar.json 2> gpurun_out/r2c5b_bench_n2_ar.err
tail -3 gpurun_out/r2c5b_bench_n2_ar.err | cut -c1-250; cut -c1-300 gpurun_out/r2c5b_bench_n2_ar.json
timeout 240 python bench.py --gpus 2 --impl reference --algo adpsgd --steps 20 --warmup 10 > gpurun_out/r2c5b_ref_n2_adpsgd.json 2> gpurun_out/r2c5b_ref_n2_adpsgd.err
grep -v "^$" gpurun_out/r2c5b_ref_n2_adpsgd.err | grep -iv "omp_num\|^\*\*\*\|warn" | tail -12 | cut -c1-250; cat gpurun_out/r2c5b_ref_n2_adpsgd.json | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments 4 2>/dev/null | grep '^{'
timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments 4 --no-pipe 2>/dev/null | grep '^{'
