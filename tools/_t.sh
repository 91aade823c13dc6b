timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "pod_by" 2>&1 | tail -3
for n in 1048576 3145728; do timeout 200 python tools/_t2.py $n 2>&1 | grep -E "n=|bias|rror" ; done
for p in 1 0; do timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --path $p 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['path'], d['pod_init'], d['pod_sigma'], d['ms_per_step'])"; done
