#!/bin/bash
# round-2 GPU session C: tests, bench arms, extra config lines, launch list, sanitizer
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; tail -3 $O/r02c_pytest.log
python __graft_entry__.py smoke > $O/r02c_smoke.log 2>&1; tail -3 $O/r02c_smoke.log
python bench.py > $O/r02c_bench.json 2> $O/r02c_bench.err; cut -c1-400 $O/r02c_bench.json
python bench.py --impl reference --steps 4 --warmup 1 > $O/r02c_bench_ref.json 2> $O/r02c_bench_ref.err; cut -c1-300 $O/r02c_bench_ref.json
python bench.py --batch 1 --steps 100 --warmup 25 --no-cpu-baseline > $O/r02c_bench_b1.json 2>$O/r02c_b1.err; cut -c1-200 $O/r02c_bench_b1.json
python bench.py --model vit_l --batch 8 --steps 20 --no-cpu-baseline > $O/r02c_bench_vitl_b8.json 2>$O/r02c_vitl.err; cut -c1-200 $O/r02c_bench_vitl_b8.json
python bench.py --bits 3 --act-order --steps 10 --no-cpu-baseline > $O/r02c_bench_int3_actorder.json 2>$O/r02c_i3.err; cut -c1-200 $O/r02c_bench_int3_actorder.json
python bench.py --bits 8 --act-order --steps 10 --no-cpu-baseline > $O/r02c_bench_int8_actorder.json 2>$O/r02c_i8.err; cut -c1-200 $O/r02c_bench_int8_actorder.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02c_ncu.log 2>&1; tail -2 $O/r02c_ncu.log | cut -c1-200
timeout 400 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > $O/r02c_memcheck.log 2>&1; tail -5 $O/r02c_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > $O/r02c_racecheck.log 2>&1; tail -5 $O/r02c_racecheck.log
