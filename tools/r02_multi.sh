#!/bin/bash
# round 2: the multi-GPU legs. usage: r02_multi.sh N [smoke]
N=$1; mode=${2:-full}
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
nvidia-smi topo -m > $out/r02_topo_n$N.txt 2>&1; lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name" >> $out/r02_topo_n$N.txt; nproc >> $out/r02_topo_n$N.txt
for i in $(seq 0 $((N-1))); do b=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $i | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "gpu $i $b numa $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null)" >> $out/r02_topo_n$N.txt; done
cat $out/r02_topo_n$N.txt | tail -16
if [ "$mode" = smoke ]; then C4N=64; C3N=2; else C4N=512; C3N=16; fi
DG=$out/r02_digests.json
if [ ! -f profiles/c4_digest.json ] || [ "$mode" = smoke ]; then
  python bench.py --workload c4 --list-videos $C4N --no-cpu --write-digest $DG > $out/r02_c4_${C4N}_n1.json 2>> $out/r02_multi.err
  python bench.py --workload c3 --list-videos $C3N --no-cpu --write-digest $DG > $out/r02_c3_${C3N}_n1.json 2>> $out/r02_multi.err
else cp profiles/c4_digest.json $DG; fi
run bench.py --gpus $N --steps 10 --warmup 3 > $out/r02_bench_c2_n$N.json 2>> $out/r02_multi.err
run bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-k1 --e2e-pool torch > $out/r02_bench_c2_n${N}_torchpool.json 2>> $out/r02_multi.err
if [ "$mode" = smoke ]; then run bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $out/r02_bench_c2_n${N}_reference.json 2>> $out/r02_multi.err; fi
run bench.py --workload c4 --list-videos $C4N --gpus $N --digest-file $DG > $out/r02_c4_${C4N}_n$N.json 2>> $out/r02_multi.err
run bench.py --workload c3 --list-videos $C3N --gpus $N --digest-file $DG > $out/r02_c3_${C3N}_n$N.json 2>> $out/r02_multi.err
if [ "$mode" = smoke ]; then run bench.py --workload c5 --gpus $N --no-cpu > $out/r02_c5_n$N.json 2>> $out/r02_multi.err; else run bench.py --workload c5 --gpus $N > $out/r02_c5_n$N.json 2>> $out/r02_multi.err; fi
tail -5 $out/r02_multi.err
for f in $out/r02_bench_c2_n$N.json $out/r02_bench_c2_n${N}_torchpool.json $out/r02_c4_${C4N}_n1.json $out/r02_c4_${C4N}_n$N.json $out/r02_c3_${C3N}_n1.json $out/r02_c3_${C3N}_n$N.json $out/r02_c5_n$N.json $out/r02_bench_c2_n${N}_reference.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    e=d.get('e2e') or {}
    print(sys.argv[1].split('/')[-1], 'value', round(d['value'],1), 'n', d.get('n_gpus'), 'e2e', round(e.get('value',0),1), 'h2d GB/s/gpu', e.get('h2d_gbs_per_gpu'), 'results', d.get('results'), 'timed', d.get('timed_region'), (d.get('cpu_baseline') or {}).get('cores'))
except Exception as ex: print(sys.argv[1], 'ERR', ex)
PY
done
