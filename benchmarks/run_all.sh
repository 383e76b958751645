#!/bin/bash
# One-GPU sweep of every benchmark of the repo; JSON lines land in gpurun_out/r01_sweep/.
#   gpurun --timeout 1500 -- 'bash benchmarks/run_all.sh'
set -u
OUT=gpurun_out/r01_sweep
mkdir -p $OUT
python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1; tail -1 $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_reference.json
python bench.py 2>/dev/null | tail -1 > $OUT/bench_c2.json
python benchmarks/beam_c3.py 2>/dev/null | tail -1 > $OUT/c3_beam.json
python benchmarks/c4_int8.py 2>/dev/null | tail -1 > $OUT/c4_kernels.json
python benchmarks/c4_decoder.py --batch 128 2>/dev/null | tail -1 > $OUT/c4_decoder_b128.json
python benchmarks/c4_decoder.py --decoder cuda --batch 64 --layers 28 2>/dev/null | tail -1 > $OUT/c2_decoder_b64.json
python benchmarks/c1_generate.py 2>/dev/null | tail -1 > $OUT/c1_generate.json
python benchmarks/gemm_i8.py 2>/dev/null | tail -1 > $OUT/gemm_m256.json
M=2048 python benchmarks/gemm_i8.py 2>/dev/null | tail -1 > $OUT/gemm_m2048.json
M=32 python benchmarks/gemm_i8.py 2>/dev/null | tail -1 > $OUT/gemm_m32.json
python benchmarks/prefill.py 2>/dev/null | tail -1 > $OUT/prefill.json
HEAD_DIM=64 python benchmarks/prefill.py 2>/dev/null | tail -1 > $OUT/prefill_d64.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 \
    benchmarks/splitkv_c5.py --ctx 16384 --iters 40 2>/dev/null | tail -1 > $OUT/c5_share_1gpu.json
ls -la $OUT
