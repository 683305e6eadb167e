L=sam_quantization_b200/lib
cp $L/libsamq.so $L/e4.so.alt
for v in e4 e0 e2 e8; do
  [ $v = e4 ] || cp $L/libsamq_$v.so.alt $L/libsamq.so
  [ $v = e4 ] && cp $L/e4.so.alt $L/libsamq.so
  echo "== $v"
  timeout 200 python -m pytest tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -1
  timeout 100 python tests/probe_gpu.py attn 2>&1 | grep "us " 
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bench', round(d['value'],2))"
done
cp $L/e4.so.alt $L/libsamq.so
