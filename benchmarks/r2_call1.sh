#!/bin/bash
# round-2 GPU call 1: state of the fp32 (precision-matched) path before any fp32 work
mkdir -p gpurun_out
python bench.py --dtype fp32 --steps 20 --warmup 5 > gpurun_out/r2c1_fp32_bs256.json 2> gpurun_out/r2c1_fp32_bs256.err
python bench.py --dtype fp32 --batch-size 32 --steps 50 --warmup 5 > gpurun_out/r2c1_fp32_bs32.json 2> gpurun_out/r2c1_fp32_bs32.err
python bench.py --dtype bf16 --batch-size 32 --steps 50 --warmup 5 > gpurun_out/r2c1_bf16_bs32.json 2> gpurun_out/r2c1_bf16_bs32.err
python bench.py --impl reference --batch-size 32 --steps 50 --warmup 10 > gpurun_out/r2c1_ref_bs32.json 2> gpurun_out/r2c1_ref_bs32.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c1_launches_fp32.csv \
   python bench.py --dtype fp32 --steps 1 --warmup 3 --no-graph --skip-e2e > gpurun_out/r2c1_ncu.log 2>&1
tail -3 gpurun_out/*.err
cat gpurun_out/r2c1_*.json
