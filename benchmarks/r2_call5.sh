#!/bin/bash
# round-2 GPU call 5 (2 GPUs): split-group pipe kernel, NVLS over unix-socket fd passing, loop-back tests, sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_gdp_loopback_gpu.py tests/test_adpsgd_gpu.py tests/test_flagship_gpu.py tests/test_conv1x1_gpu.py -q > gpurun_out/r2c5_tests_1gpu.log 2>&1
tail -15 gpurun_out/r2c5_tests_1gpu.log
timeout 1200 python -m pytest tests/test_multigpu.py -q > gpurun_out/r2c5_tests_multigpu_n2.log 2>&1
tail -25 gpurun_out/r2c5_tests_multigpu_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
for seg in 4 7 12; do
  timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments $seg 2>/dev/null | grep '^{' >> gpurun_out/r2c5_mix_bench.jsonl
done
timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments 4 --no-pipe 2>/dev/null | grep '^{' >> gpurun_out/r2c5_mix_bench.jsonl
timeout 120 $TR benchmarks/mix_bench.py --mode local 2>/dev/null | grep '^{' >> gpurun_out/r2c5_mix_bench.jsonl
timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments 7 --ppi 1 --bf16 2>/dev/null | grep '^{' >> gpurun_out/r2c5_mix_bench.jsonl
cat gpurun_out/r2c5_mix_bench.jsonl
timeout 120 python benchmarks/profile_mix.py --two-gpu --iters 12 > gpurun_out/r2c5_two_gpu_mix.log 2>&1; cat gpurun_out/r2c5_two_gpu_mix.log
timeout 400 ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --clock-control none --import-source on \
   -k regex:sgp_step_pipe -s 3 -c 2 -o gpurun_out/r2c5_prof_step_pipe_2gpu python benchmarks/profile_mix.py --two-gpu --iters 6 > gpurun_out/r2c5_ncu_pipe.log 2>&1
tail -2 gpurun_out/r2c5_ncu_pipe.log
for algo in sgp osgp adpsgd ar; do
  timeout 300 python bench.py --gpus 2 --algo $algo --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c5_bench_n2_$algo.json 2> gpurun_out/r2c5_bench_n2_$algo.err
  tail -2 gpurun_out/r2c5_bench_n2_$algo.err | cut -c1-300; cut -c1-300 gpurun_out/r2c5_bench_n2_$algo.json
done
for tr in p2p nccl; do
  timeout 300 python bench.py --gpus 2 --algo ar --ar-transport $tr --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c5_bench_n2_ar_$tr.json 2> gpurun_out/r2c5_bench_n2_ar_$tr.err
  cut -c1-300 gpurun_out/r2c5_bench_n2_ar_$tr.json
done
timeout 400 python bench.py --gpus 2 --impl reference --algo adpsgd --steps 20 --warmup 10 > gpurun_out/r2c5_ref_n2_adpsgd.json 2> gpurun_out/r2c5_ref_n2_adpsgd.err
tail -5 gpurun_out/r2c5_ref_n2_adpsgd.err | cut -c1-300; cat gpurun_out/r2c5_ref_n2_adpsgd.json | cut -c1-500
for algo in sgp osgp; do
  timeout 300 $TR benchmarks/e2e_convergence.py --algo $algo --iters 200 --out gpurun_out/r2c5_e2e_n2_$algo.json > gpurun_out/r2c5_e2e_n2_$algo.log 2>&1
  tail -1 gpurun_out/r2c5_e2e_n2_$algo.log | cut -c1-600
done
CUDA_VISIBLE_DEVICES=0 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "fused_sgd_mix_matches_oracle and pipe" > gpurun_out/r2c5_sanitizer_racecheck_step_pipe.log 2>&1
tail -8 gpurun_out/r2c5_sanitizer_racecheck_step_pipe.log | cut -c1-200
CUDA_VISIBLE_DEVICES=0 timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "fused_sgd_mix_matches_oracle and pipe" > gpurun_out/r2c5_sanitizer_synccheck_step_pipe.log 2>&1
tail -8 gpurun_out/r2c5_sanitizer_synccheck_step_pipe.log | cut -c1-200
