#!/bin/bash
# tools/r2_mg.sh N TAG: the driver's multi-GPU launches of bench.py (C2 + extra, then the C4 / C5 headline lines, then the reference arm)
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-r2mg}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
mkdir -p gpurun_out
$TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${TAG}_c2_$N.json 2> gpurun_out/${TAG}_c2_$N.err || tail -5 gpurun_out/${TAG}_c2_$N.err
$TR bench.py --gpus $N --workload c4 --steps 20 --warmup 3 > gpurun_out/${TAG}_c4_$N.json 2> gpurun_out/${TAG}_c4_$N.err || tail -5 gpurun_out/${TAG}_c4_$N.err
$TR bench.py --gpus $N --workload c5 --steps 10 --warmup 3 > gpurun_out/${TAG}_c5_$N.json 2> gpurun_out/${TAG}_c5_$N.err || tail -5 gpurun_out/${TAG}_c5_$N.err
$TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_$N.json 2> gpurun_out/${TAG}_ref_$N.err || tail -5 gpurun_out/${TAG}_ref_$N.err
python - <<PY
import json
def last(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        return {'error': repr(e)}
d=last("gpurun_out/${TAG}_c2_$N.json")
print('C2', d.get('n_gpus'), d.get('value'), d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('ms_per_step_by_frames_in_flight'), (d.get('e2e') or {}).get('d2h_only_ms'))
print('extra', d.get('extra'))
for k in ('c4','c5','ref'):
    d=last("gpurun_out/${TAG}_%s_$N.json" % k)
    print(k, d.get('value'), d.get('unit'), d.get('ms_per_step'), d.get('detail') or d.get('cpu_baseline') or d.get('error'))
PY
