#!/bin/bash
# tools/run_multigpu.sh N  -- the multi-GPU measurements of SURVEY §8e on one node (bench C2 weak scaling,
# C4 view-parallel, C5 sort-last with the bit-identity checksum against the 1-GPU result)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
$TR bench.py --gpus $N --steps 100 --warmup 10 2>gpurun_out/mg_bench_$N.err | tail -1 > gpurun_out/mg_bench_$N.json
$TR tools/bench_configs.py c4 2>/dev/null | tail -1 > gpurun_out/mg_c4_$N.json
$TR tools/bench_configs.py c5 --check --iters 5 2>/dev/null | tail -1 > gpurun_out/mg_c5_$N.json
$TR tools/bench_configs.py c5 --check --iters 5 --partitioned 2>/dev/null | tail -1 > gpurun_out/mg_c5p_$N.json
python tools/bench_configs.py c5 --check --iters 3 2>/dev/null | tail -1 > gpurun_out/mg_c5_1.json
for f in mg_bench_$N mg_c4_$N mg_c5_$N mg_c5p_$N mg_c5_1; do echo "== $f"; cut -c1-700 gpurun_out/$f.json; done
