#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_allreduce_check.py > $OUT/ddp_allreduce_n2.json 2> $OUT/ddp_allreduce_n2.err; echo "ddp rc=$?"; tail -2 $OUT/ddp_allreduce_n2.json; tail -3 $OUT/ddp_allreduce_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 6 --warmup 3 --no-extra > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"; python -c "
import json;d=json.loads(open('$OUT/bench_n2.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])"
