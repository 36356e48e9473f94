#!/bin/bash
# round-2 GPU call 4 (2 GPUs): AD-PSGD daemon, NVLS, single-process replicas, pipe kernel at 2 CTAs/SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_adpsgd_gpu.py tests/test_conv1x1_gpu.py tests/test_flagship_gpu.py tests/test_kernels_gpu.py tests/test_fused_bn_gpu.py tests/test_fused_loss_gpu.py -q -x > gpurun_out/r2c4_tests_1gpu.log 2>&1
tail -15 gpurun_out/r2c4_tests_1gpu.log
timeout 1200 python -m pytest tests/test_multigpu.py -q > gpurun_out/r2c4_tests_multigpu_n2.log 2>&1
tail -25 gpurun_out/r2c4_tests_multigpu_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
for seg in 4 7; do
  timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments $seg 2>/dev/null | grep '^{' >> gpurun_out/r2c4_mix_bench.jsonl
done
timeout 120 $TR benchmarks/mix_bench.py --mode mix --segments 4 --no-pipe 2>/dev/null | grep '^{' >> gpurun_out/r2c4_mix_bench.jsonl
timeout 120 $TR benchmarks/mix_bench.py --mode local 2>/dev/null | grep '^{' >> gpurun_out/r2c4_mix_bench.jsonl
timeout 120 $TR benchmarks/mix_bench.py --mode publish 2>/dev/null | grep '^{' >> gpurun_out/r2c4_mix_bench.jsonl
timeout 120 $TR benchmarks/mix_bench.py --mode gather 2>/dev/null | grep '^{' >> gpurun_out/r2c4_mix_bench.jsonl
cat gpurun_out/r2c4_mix_bench.jsonl
timeout 120 python benchmarks/profile_mix.py --two-gpu --iters 12 > gpurun_out/r2c4_two_gpu_mix.log 2>&1; cat gpurun_out/r2c4_two_gpu_mix.log
timeout 400 ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --clock-control none --import-source on \
   -k regex:sgp_step_pipe -s 3 -c 2 -o gpurun_out/r2c4_prof_step_pipe_2gpu python benchmarks/profile_mix.py --two-gpu --iters 6 > gpurun_out/r2c4_ncu_pipe.log 2>&1
tail -2 gpurun_out/r2c4_ncu_pipe.log
for algo in sgp osgp adpsgd ar; do
  timeout 300 python bench.py --gpus 2 --algo $algo --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c4_bench_n2_$algo.json 2> gpurun_out/r2c4_bench_n2_$algo.err
  tail -2 gpurun_out/r2c4_bench_n2_$algo.err | cut -c1-300; cut -c1-400 gpurun_out/r2c4_bench_n2_$algo.json
done
timeout 300 python bench.py --gpus 2 --algo ar --ar-transport p2p --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c4_bench_n2_ar_p2p.json 2> gpurun_out/r2c4_bench_n2_ar_p2p.err
cut -c1-300 gpurun_out/r2c4_bench_n2_ar_p2p.json
timeout 400 python bench.py --gpus 2 --impl reference --algo adpsgd --steps 20 --warmup 10 > gpurun_out/r2c4_ref_n2_adpsgd.json 2> gpurun_out/r2c4_ref_n2_adpsgd.err
tail -5 gpurun_out/r2c4_ref_n2_adpsgd.err | cut -c1-300; cat gpurun_out/r2c4_ref_n2_adpsgd.json | cut -c1-500
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c4_bench_n1.json 2> gpurun_out/r2c4_bench_n1.err
tail -2 gpurun_out/r2c4_bench_n1.err | cut -c1-300; cat gpurun_out/r2c4_bench_n1.json
