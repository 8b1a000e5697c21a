#!/bin/bash
# round-2 pass i (gpurun --gpus 2): multi-rank parity with the collective reaction pass, configs[2] at 200^3 on 1 and 2 GPUs
N=${1:-2}; TAG=${2:-r3i}
O=gpurun_out/$TAG
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 240 $TR --master-port 29514 tests/multigpu_check.py > $O/check_n$N.log 2>&1; echo "check exit $?" >> $O/check_n$N.log
grep -E "reaction|support|exit|rror" $O/check_n$N.log | tail -8; grep -c " OK" $O/check_n$N.log
timeout -k 10 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29516 tools/bench_pml_multi.py --size 200 --steps 20 > $O/bench_pml_n1.json 2> $O/bench_pml_n1.err
timeout -k 10 330 $TR --master-port 29515 tools/bench_pml_multi.py --size 200 --steps 20 > $O/bench_pml_n$N.json 2> $O/bench_pml_n$N.err
cat $O/bench_pml_n1.json $O/bench_pml_n$N.json; tail -2 $O/bench_pml_n$N.err
