#!/bin/bash
# round-2 GPU session H: full suite, reference-on-GPU harness (microbench / attention / encoder), both bench arms as the driver runs them
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02h_pytest.log 2>&1; tail -3 $O/r02h_pytest.log
rm -f $O/ref_gpu_r02h.json
python oracle/ref_gpu.py --sections dequant,microbench,attention,encoder --out $O/ref_gpu_r02h > $O/r02h_refgpu.log 2>&1; tail -2 $O/r02h_refgpu.log | cut -c1-300
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02h_bench_ref.json 2> $O/r02h_bench_ref.err; cut -c1-200 $O/r02h_bench_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02h_bench.json 2> $O/r02h_bench.err; cut -c1-200 $O/r02h_bench.json
for b in 8 16 64; do python bench.py --batch $b --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch', $b, round(d['value'],1), round(d['e2e']['value'],1))"; done > $O/r02h_batch_sweep.log; cat $O/r02h_batch_sweep.log
