#!/bin/bash
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
F=${2:-33554432}
$TR tools/bench_configs.py c5 --check --iters 5 --faces $F 2>&1 | tail -1
$TR tools/bench_configs.py c5 --check --iters 5 --faces $F --p2p 2>&1 | tail -1
$TR tools/bench_configs.py c5 --check --iters 5 --faces $F --p2p --root 2>&1 | tail -3
