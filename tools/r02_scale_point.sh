#!/bin/bash
# one more point of the scaling curves: C4 (strong) and the C2 e2e leg at N GPUs
N=$1; out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
run bench.py --workload c4 --gpus $N --no-cpu > $out/r02_c4_512_n$N.json 2>> $out/r02_multi.err
if [ "$2" = e2e ]; then run bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-k1 > $out/r02_bench_c2_n$N.json 2>> $out/r02_multi.err; fi
for f in $out/r02_c4_512_n$N.json $out/r02_bench_c2_n$N.json; do [ -f $f ] && python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
e=d.get('e2e') or {}
print(sys.argv[1].split('/')[-1], round(d['value'],1), 'n', d.get('n_gpus'), 'e2e', e.get('value'), e.get('h2d_gbs_per_gpu'), e.get('h2d_probe_gbs_per_gpu'), d.get('results'), {k:v for k,v in (d.get('timed_region') or {}).items() if k!='note'})
PY
done
