#!/bin/bash
# round-2 GPU call 3 (2 GPUs): NVLS probe, 1-GPU tests of the new fp32 path, multi-GPU tests,
# step-kernel A/B over NVLink, ncu with NVLink counters, SGP / OSGP exposed comm
mkdir -p gpurun_out
./benchmarks/nvls/nvls_probe > gpurun_out/r2c3_nvls_probe.log 2>&1; cat gpurun_out/r2c3_nvls_probe.log
timeout 600 python -m pytest tests/test_conv1x1_gpu.py tests/test_fused_loss_gpu.py tests/test_fused_bn_gpu.py tests/test_flagship_gpu.py tests/test_kernels_gpu.py -q > gpurun_out/r2c3_tests_1gpu.log 2>&1
tail -12 gpurun_out/r2c3_tests_1gpu.log
timeout 900 python -m pytest tests/test_multigpu.py -q > gpurun_out/r2c3_tests_multigpu_n2.log 2>&1
tail -12 gpurun_out/r2c3_tests_multigpu_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for seg in 4 7 12; do
  timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments $seg 2>/dev/null | grep '^{' >> gpurun_out/r2c3_mix_bench.jsonl
done
timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments 4 --no-pipe 2>/dev/null | grep '^{' >> gpurun_out/r2c3_mix_bench.jsonl
timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments 7 --ppi 1 --bf16 2>/dev/null | grep '^{' >> gpurun_out/r2c3_mix_bench.jsonl
timeout 120 $TR benchmarks/mix_bench.py --mode local 2>/dev/null | grep '^{' >> gpurun_out/r2c3_mix_bench.jsonl
cat gpurun_out/r2c3_mix_bench.jsonl
timeout 120 python benchmarks/profile_mix.py --two-gpu --iters 12 > gpurun_out/r2c3_two_gpu_mix.log 2>&1; cat gpurun_out/r2c3_two_gpu_mix.log
timeout 120 python benchmarks/profile_mix.py --two-gpu --iters 12 --no-pipe >> gpurun_out/r2c3_two_gpu_mix.log 2>&1; tail -1 gpurun_out/r2c3_two_gpu_mix.log
timeout 400 ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --clock-control none --import-source on \
   -k regex:sgp_step_pipe -s 3 -c 2 -o gpurun_out/r2c3_prof_step_pipe_2gpu python benchmarks/profile_mix.py --two-gpu --iters 6 > gpurun_out/r2c3_ncu_pipe.log 2>&1
tail -3 gpurun_out/r2c3_ncu_pipe.log
for algo in sgp osgp; do
  timeout 300 python bench.py --gpus 2 --algo $algo --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c3_bench_n2_$algo.json 2> gpurun_out/r2c3_bench_n2_$algo.err
  tail -2 gpurun_out/r2c3_bench_n2_$algo.err; cat gpurun_out/r2c3_bench_n2_$algo.json
done
timeout 200 $TR benchmarks/osgp_trace.py --algo osgp --batch-size 64 --out gpurun_out/r2c3_osgp_trace.json > gpurun_out/r2c3_osgp_trace.log 2>&1; tail -8 gpurun_out/r2c3_osgp_trace.log
timeout 200 $TR benchmarks/osgp_trace.py --algo sgp --batch-size 64 > gpurun_out/r2c3_sgp_trace.log 2>&1; tail -6 gpurun_out/r2c3_sgp_trace.log
