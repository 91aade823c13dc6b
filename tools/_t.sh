timeout 100 python tools/_t.py 151552 1000 2>&1 | grep -E "^OK|FAILED|stuck|progress" | head -12
timeout 120 python bench.py --steps 10 --warmup 3 --path 2 --no-cpu --no-e2e 2> gpurun_out/bench3.err | tail -1 > gpurun_out/bench3.json; python -c "
import json; d=json.load(open('gpurun_out/bench3.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
timeout 100 python tools/tc_timers.py 2>&1 | grep -E "kernel|mma|epi:|prod" | head -24
