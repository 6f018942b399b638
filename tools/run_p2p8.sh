#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
$TR tools/bench_configs.py c5 --check --iters 5 --p2p 2>/dev/null | tail -1 | tee gpurun_out/mg_c5_p2p_$N.json
$TR tools/bench_configs.py c5 --check --iters 5 --p2p --root 2>/dev/null | tail -1 | tee gpurun_out/mg_c5_p2p_root_$N.json
$TR tools/bench_configs.py c5 --check --iters 5 --root 2>/dev/null | tail -1 | tee gpurun_out/mg_c5_rs_root_$N.json
$TR tools/bench_configs.py c4 --graph 2>/dev/null | tail -1 | tee gpurun_out/mg_c4_graph_$N.json
