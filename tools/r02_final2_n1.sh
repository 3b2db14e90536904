#!/bin/bash
# round 2, last single-GPU evidence after the one-warp-per-tile votes kernel: the whole GPU suite on the product build, smoke,
# the default bench line (C2) and C1.
out=gpurun_out; mkdir -p $out
timeout 270 python -m pytest tests -m gpu -x -q > $out/r02_gputest_product2.log 2>&1; echo "exit $?" >> $out/r02_gputest_product2.log; tail -3 $out/r02_gputest_product2.log
timeout 45 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 100 python bench.py --steps 20 > $out/r02_bench_c2_final2.json 2> $out/r02_bench_c2_final2.err; echo "c2 exit $?"; tail -c 300 $out/r02_bench_c2_final2.err
timeout 60 python bench.py --workload c1 --steps 50 --no-cpu > $out/r02_bench_c1_final2.json 2>> $out/r02_bench_c2_final2.err; echo "c1 exit $?"
for f in r02_bench_c2_final2 r02_bench_c1_final2; do python - $out/$f.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
except Exception as e:
    print(sys.argv[1], 'no line', e); sys.exit(0)
r=d.get('roofline') or {}
print(sys.argv[1].split('/')[-1], round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'k2', round(r.get('frac',0),4), 'traffic', r.get('traffic'), 'e2e', (d.get('e2e') or {}).get('value'), d.get('stage_ms'), (d.get('results') or {}), (d.get('overlap_gemm') or {}).get('frac'))
PY
done
