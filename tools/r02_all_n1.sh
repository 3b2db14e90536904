#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r02_gputest_product.log 2>&1; echo "exit $?" >> $out/r02_gputest_product.log; tail -3 $out/r02_gputest_product.log
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_check.so python -m pytest tests -m gpu -q > $out/r02_gputest_boundscheck.log 2>&1; echo "exit $?" >> $out/r02_gputest_boundscheck.log; tail -3 $out/r02_gputest_boundscheck.log
python tools/k1_bench.py > $out/r02_k1_gram_bench_v6.json 2> $out/r02_k1_bench.err; python -c "
import json; a=json.load(open('$out/r02_k1_gram_bench_v6.json'))
for k,v in a.items(): print(k, v if not isinstance(v,dict) else (round(v['ms_per_video'],4), round(v['frac_of_measured_int8_peak'],3)))"
python bench.py --no-e2e --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['stage_ms'], d['parity_check']); print(json.dumps(d['overlap_gemm'])[:900])"
python bench.py --workload c1 --no-e2e --no-cpu --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c1', round(d['value']), d['ms_per_step'], d['stage_ms']); print(json.dumps(d['overlap_gemm'])[:900])"
