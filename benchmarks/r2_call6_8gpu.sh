#!/bin/bash
# round-2 GPU call 6 (8 GPUs): every BASELINE config at N=8 (ours + reference arms), multi-GPU tests,
# gossip kernel at ppi 1/2, end-to-end convergence tier
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 300 python -m pytest tests/test_multigpu.py -q -k "nvls_allreduce or hierarchical or adpsgd_device or random_skew or sgp_mix" > gpurun_out/r2c6_tests_multigpu_n8.log 2>&1
tail -6 gpurun_out/r2c6_tests_multigpu_n8.log | cut -c1-200
for algo in sgp osgp dpsgd adpsgd ar; do
  timeout 150 python bench.py --gpus $N --algo $algo --steps 12 --warmup 5 --no-secondary > gpurun_out/r2c6_bench_n8_$algo.json 2> gpurun_out/r2c6_bench_n8_$algo.err
  tail -1 gpurun_out/r2c6_bench_n8_$algo.err | cut -c1-200; cut -c1-220 gpurun_out/r2c6_bench_n8_$algo.json
done
for tr in p2p nccl; do
  timeout 150 python bench.py --gpus $N --algo ar --ar-transport $tr --steps 12 --warmup 5 --no-secondary --skip-e2e > gpurun_out/r2c6_bench_n8_ar_$tr.json 2> gpurun_out/r2c6_bench_n8_ar_$tr.err
  cut -c1-220 gpurun_out/r2c6_bench_n8_ar_$tr.json
done
timeout 150 python bench.py --gpus $N --algo sgp --batch-size 32 --steps 30 --warmup 5 --no-secondary > gpurun_out/r2c6_bench_n8_sgp_bs32.json 2> gpurun_out/r2c6_bench_n8_sgp_bs32.err
cut -c1-220 gpurun_out/r2c6_bench_n8_sgp_bs32.json
for algo in osgp dpsgd ar; do
  timeout 200 python bench.py --gpus $N --impl reference --algo $algo --steps 12 --warmup 6 > gpurun_out/r2c6_ref_n8_$algo.json 2> gpurun_out/r2c6_ref_n8_$algo.err
  tail -1 gpurun_out/r2c6_ref_n8_$algo.err | cut -c1-200; cut -c1-220 gpurun_out/r2c6_ref_n8_$algo.json
done
for ppi in 1 2; do
  timeout 60 $TR benchmarks/mix_bench.py --mode mix --ppi $ppi 2>/dev/null | grep '^{' >> gpurun_out/r2c6_mix_bench_n8.jsonl
done
cat gpurun_out/r2c6_mix_bench_n8.jsonl
timeout 150 $TR benchmarks/e2e_convergence.py --algo sgp --iters 200 --out gpurun_out/r2c6_e2e_n8_sgp.json > gpurun_out/r2c6_e2e_n8_sgp.log 2>&1
tail -1 gpurun_out/r2c6_e2e_n8_sgp.log | cut -c1-400
