timeout 100 python tools/tc_timers.py
timeout 120 python bench.py --steps 10 --warmup 3 --path 2 --no-cpu --no-e2e 2> gpurun_out/bench3.err | tail -1 > gpurun_out/bench3.json; python -c "
import json; d=json.load(open('gpurun_out/bench3.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
