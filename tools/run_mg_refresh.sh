#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
$TR bench.py --gpus $N --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/mg2_bench_$N.json; cut -c1-200 gpurun_out/mg2_bench_$N.json
$TR tools/bench_configs.py c5 --check --iters 5 2>/dev/null | tail -1 | tee gpurun_out/mg2_c5_$N.json
$TR tools/bench_configs.py c5 --check --iters 5 --p2p --root 2>/dev/null | tail -1 | tee gpurun_out/mg2_c5_p2p_root_$N.json
$TR tools/bench_configs.py c4 --graph 2>/dev/null | tail -1 | tee gpurun_out/mg2_c4_graph_$N.json
