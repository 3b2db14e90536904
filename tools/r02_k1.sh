#!/bin/bash
# round-2 K1 (Gram overlap) check + timing + ncu on the GPU box
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram or overlap" > $out/r02_k1_tests.log 2>&1; echo "exit $?" >> $out/r02_k1_tests.log; tail -3 $out/r02_k1_tests.log
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_check.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gram or overlap" > $out/r02_k1_tests_check.log 2>&1; echo "exit $?" >> $out/r02_k1_tests_check.log; tail -3 $out/r02_k1_tests_check.log
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/k1_bench.py > $out/r02_k1_gram_bench_v5_zero_skip.json 2> $out/r02_k1_bench.err; cat $out/r02_k1_gram_bench_v5_zero_skip.json | head -c 3000
python bench.py --no-e2e --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(d['overlap_gemm'])[:1500])"
python tools/k1_one.py > /dev/null
ncu --set full --clock-control none --import-source on -k regex:gram_labels2 -c 1 -f -o $out/r02_gram_v5 python tools/k1_one.py > $out/ncu_gram5.log 2>&1
ls -la $out/r02_gram_v5.ncu-rep
