#!/bin/bash
# round-2 GPU session T (final tree): launch list of the default bench command + ncu --set full of the
# LayerNorm (three resident blocks per SM) and attention kernels (bias gather / O drain with one wait)
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/r02t_nograph.json 2> $O/r02t_nograph.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/r02t_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/r02t_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"layernorm|attn_" --launch-skip 40 --launch-count 10 -f -o $O/r02t_ln_attn python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02t_ln_attn.log 2>&1
tail -1 $O/r02t_ln_attn.log | cut -c1-160
ls -la $O/r02t*
