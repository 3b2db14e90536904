#!/bin/bash
# round 2: the list workloads at full size on one GPU (reference digests for the multi-GPU runs) + K1 re-check
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or overlap" > $out/r02_k1_tests.log 2>&1; echo "exit $?" >> $out/r02_k1_tests.log; tail -2 $out/r02_k1_tests.log
python tools/k1_bench.py > $out/r02_k1_gram_bench_v6.json 2> $out/r02_k1_bench.err; python -c "
import json; a=json.load(open('$out/r02_k1_gram_bench_v6.json'))
for k,v in a.items(): print(k, v if not isinstance(v,dict) else (round(v['ms_per_video'],4), round(v['frac_of_measured_int8_peak'],3)))"
(time python bench.py --workload c4 > $out/r02_c4_n1.json) 2> $out/r02_c4_n1.err; tail -4 $out/r02_c4_n1.err; head -c 2500 $out/r02_c4_n1.json; echo
(time python bench.py --workload c3 > $out/r02_c3_n1.json) 2> $out/r02_c3_n1.err; tail -4 $out/r02_c3_n1.err; head -c 2500 $out/r02_c3_n1.json; echo
(time python bench.py --workload c5 > $out/r02_c5_n1.json) 2> $out/r02_c5_n1.err; tail -4 $out/r02_c5_n1.err; head -c 1500 $out/r02_c5_n1.json; echo
