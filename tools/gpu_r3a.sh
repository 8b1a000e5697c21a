#!/bin/bash
# round-2 first pass (gpurun --gpus 2): everything that never ran on hardware, bounded timeouts
N=${1:-2}; TAG=${2:-r3a}
O=gpurun_out/$TAG
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 420 $TR --master-port 29514 tests/multigpu_check.py > $O/check_pml_n$N.log 2>&1; echo "check(pml) exit $?" >> $O/check_pml_n$N.log
grep -E "multigpu|exit|rror" $O/check_pml_n$N.log | tail -30
timeout -k 10 300 python tests/multigpu_host_check.py $N > $O/check_host_n$N.log 2>&1; echo "check(host) exit $?" >> $O/check_host_n$N.log
grep -E "multigpu host|exit" $O/check_host_n$N.log | tail -8
if grep -q "check(pml) exit 0" $O/check_pml_n$N.log; then
  timeout -k 10 400 $TR --master-port 29515 tools/bench_pml_multi.py --size ${PML_N:-200} --steps 20 > $O/bench_pml_n$N.json 2> $O/bench_pml_n$N.err
  cat $O/bench_pml_n$N.json; tail -3 $O/bench_pml_n$N.err
fi
