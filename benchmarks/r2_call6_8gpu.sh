#!/bin/bash
# round-2 GPU call 6 (8 GPUs): every BASELINE config at N=8 (ours + reference arms), multi-GPU tests,
# gossip kernel at ppi 1/2, end-to-end convergence tier
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 python -m pytest tests/test_multigpu.py -q -k "nvls or vmm or sgp_mix or kernels_match_simulation or adpsgd_device or random_skew or hierarchical or bounded_staleness or barrier" > gpurun_out/r2c6_tests_multigpu_n8.log 2>&1
tail -12 gpurun_out/r2c6_tests_multigpu_n8.log
for algo in sgp osgp dpsgd adpsgd ar; do
  timeout 240 python bench.py --gpus $N --algo $algo --steps 15 --warmup 5 --no-secondary > gpurun_out/r2c6_bench_n8_$algo.json 2> gpurun_out/r2c6_bench_n8_$algo.err
  tail -1 gpurun_out/r2c6_bench_n8_$algo.err | cut -c1-200; cut -c1-260 gpurun_out/r2c6_bench_n8_$algo.json
done
for tr in p2p nccl; do
  timeout 240 python bench.py --gpus $N --algo ar --ar-transport $tr --steps 15 --warmup 5 --no-secondary --skip-e2e > gpurun_out/r2c6_bench_n8_ar_$tr.json 2> gpurun_out/r2c6_bench_n8_ar_$tr.err
  cut -c1-260 gpurun_out/r2c6_bench_n8_ar_$tr.json
done
timeout 240 python bench.py --gpus $N --algo sgp --batch-size 32 --steps 30 --warmup 5 --no-secondary > gpurun_out/r2c6_bench_n8_sgp_bs32.json 2> gpurun_out/r2c6_bench_n8_sgp_bs32.err
cut -c1-260 gpurun_out/r2c6_bench_n8_sgp_bs32.json
timeout 240 python bench.py --gpus $N --algo osgp --batch-size 32 --steps 30 --warmup 5 --no-secondary > gpurun_out/r2c6_bench_n8_osgp_bs32.json 2> gpurun_out/r2c6_bench_n8_osgp_bs32.err
cut -c1-260 gpurun_out/r2c6_bench_n8_osgp_bs32.json
for algo in osgp dpsgd ar adpsgd; do
  timeout 300 python bench.py --gpus $N --impl reference --algo $algo --steps 15 --warmup 8 > gpurun_out/r2c6_ref_n8_$algo.json 2> gpurun_out/r2c6_ref_n8_$algo.err
  tail -2 gpurun_out/r2c6_ref_n8_$algo.err | cut -c1-200; cut -c1-260 gpurun_out/r2c6_ref_n8_$algo.json
done
timeout 240 python bench.py --gpus $N --impl reference --algo sgp --batch-size 32 --steps 30 --warmup 8 > gpurun_out/r2c6_ref_n8_sgp_bs32.json 2> gpurun_out/r2c6_ref_n8_sgp_bs32.err
cut -c1-260 gpurun_out/r2c6_ref_n8_sgp_bs32.json
for ppi in 1 2; do
  timeout 100 $TR benchmarks/mix_bench.py --mode mix --ppi $ppi 2>/dev/null | grep '^{' >> gpurun_out/r2c6_mix_bench_n8.jsonl
done
timeout 100 $TR benchmarks/mix_bench.py --mode mix --ppi 1 --no-pipe 2>/dev/null | grep '^{' >> gpurun_out/r2c6_mix_bench_n8.jsonl
cat gpurun_out/r2c6_mix_bench_n8.jsonl
for algo in sgp osgp; do
  timeout 240 $TR benchmarks/e2e_convergence.py --algo $algo --iters 300 --out gpurun_out/r2c6_e2e_n8_$algo.json > gpurun_out/r2c6_e2e_n8_$algo.log 2>&1
  tail -1 gpurun_out/r2c6_e2e_n8_$algo.log | cut -c1-500
done
