#!/bin/bash
# scaling curves: weak (3*2^20 points/GPU) and strong (3*2^20 and 3*2^23 points total) at N = $1
N=$1
mkdir -p gpurun_out
run() { # name, extra args
  if [ "$N" = "1" ]; then timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu --no-e2e $2 2>gpurun_out/scale_$1_n$N.err | tail -1 > gpurun_out/scale_$1_n$N.json
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-e2e $2 2>gpurun_out/scale_$1_n$N.err | tail -1 > gpurun_out/scale_$1_n$N.json; fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_$1_n$N.json')); print('$1 N=$N', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'points/gpu', d['config']['points_per_gpu'], 'parity', d.get('sharded_parity'), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
}
run weak ""
run strong20 "--scaling strong --total-points $((3*1048576))"
[ -z "$SKIP23" ] && run strong23 "--scaling strong --total-points $((3*8388608))"
