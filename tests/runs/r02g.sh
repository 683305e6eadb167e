#!/bin/bash
# round-2 GPU session G: decoder tests + ncu --set full captures of every product kernel class (final r02 tree)
O=gpurun_out
python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -3
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:dense2_kernel --launch-skip 200 --launch-count 4 -f -o $O/r02g_dense2 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02g_dense2.log 2>&1; tail -1 $O/r02g_dense2.log | cut -c1-120
$NCU -k regex:attn_ --launch-skip 20 --launch-count 8 -f -o $O/r02g_attn python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02g_attn.log 2>&1; tail -1 $O/r02g_attn.log | cut -c1-120
$NCU -k regex:"layernorm|fill_pad|dequant4" --launch-skip 60 --launch-count 6 -f -o $O/r02g_small python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02g_small.log 2>&1; tail -1 $O/r02g_small.log | cut -c1-120
SAMQ_GEMM=fused $NCU -k regex:qlinear_kernel --launch-skip 3 --launch-count 1 -f -o $O/r02g_fused_m4096 python tests/gemm_bench.py 4096 1280 5120 none 3 > $O/r02g_fused1.log 2>&1; tail -1 $O/r02g_fused1.log | cut -c1-120
$NCU -k regex:qlinear_kernel --launch-skip 3 --launch-count 1 -f -o $O/r02g_fused_m196 python tests/gemm_bench.py 196 1280 5120 none 3 > $O/r02g_fused2.log 2>&1; tail -1 $O/r02g_fused2.log | cut -c1-120
for M in 196 512 1024 2048 4096; do for v in fused dense; do SAMQ_GEMM=$v python tests/gemm_bench.py $M 1280 3840 none 50; done; done > $O/r02g_crossover.log 2>&1; cat $O/r02g_crossover.log
ls -la $O/*.ncu-rep | tail -6
