#!/bin/bash
# multi-GPU bench of library variants: run_mgpu.sh N lib [lib ...]
N=$1; shift
for v in "$@"; do
  echo -n "N=$N $v: "
  SMB200_LIB=$PWD/scripts/ab/$v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 10000 --warmup 1000 2>/dev/null > gpurun_out/ab_${N}_$v.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/ab_${N}_$v.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
