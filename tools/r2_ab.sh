#!/bin/bash
# A/B of library variants on the C2 bench: tools/r2_ab.sh tag variant1 variant2 ...  (build_variants/lib_<variant>.so)
cd "$(dirname "$0")/.."
TAG=$1; shift
for v in "$@"; do
  TINA_B200_LIB=$PWD/build_variants/lib_$v.so python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/${TAG}_ab_$v.json 2>gpurun_out/${TAG}_ab_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_ab_$v.json'))
print('$v', 'ms/step', round(d['ms_per_step']*1e3,2), 'us  kernels', {k: round(x*1e3,2) for k,x in d['kernel_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3))
PY
done
