#!/bin/bash
# scripts/run_scaling.sh N [port]  -- bench.py on N GPUs of one box, both workloads, JSON into gpurun_out/
N=$1; PORT=${2:-29720}
mkdir -p gpurun_out
run() { # name, extra args
  if [ "$N" = 1 ]; then timeout 600 python bench.py --gpus 1 $2 > gpurun_out/scale_$1_n$N.json 2> gpurun_out/scale_$1_n$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N $2 > gpurun_out/scale_$1_n$N.json 2> gpurun_out/scale_$1_n$N.err; fi
  tail -1 gpurun_out/scale_$1_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); f=d.get('fit') or {}
    print('$1 N=$N value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3), 'fit_wall', f.get('wall_s'), 'clocks', d['clocks'])
except Exception as e: print('$1 N=$N ERR', e)"
}
run cfg2 "--steps 20 --warmup 3 --skip-cpu --skip-fit --skip-large"
PORT=$((PORT+1))
run cfg4 "--steps 5 --warmup 3 --workload cfg4"
