#!/bin/bash
# round-2 A/B of the votes kernel (K2) on the GPU box: correctness first, then timings of the C2 / C1 / mixed shapes.
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_workloads.py -m gpu -x -q -k "point_votes or ragged or windowed or c1_shape or c2_shape or random_scenes or device_runner or c3_shaped or golden" > $out/r02_k2_tests.log 2>&1; echo "exit $?" >> $out/r02_k2_tests.log; tail -3 $out/r02_k2_tests.log
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_check.so python -m pytest tests -m gpu -q > $out/r02_gputest_boundscheck.log 2>&1; echo "exit $?" >> $out/r02_gputest_boundscheck.log; tail -3 $out/r02_gputest_boundscheck.log
b() { python bench.py --no-e2e --no-cpu --no-k1 "$@" 2>>$out/r02_k2_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$*', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'k2 frac', round(r['frac'],4), 'k2 ms', round(r.get('ms_per_launch', r.get('ms_sum_over_ranks',0)),3), d.get('parity_check'), {k: round(v,3) for k,v in d.get('stage_ms',{}).items()})"; }
b --steps 10
b --steps 10 --point-order random
b --workload target --steps 10
b --workload c1 --steps 50
b --workload c4 --list-videos 32
export S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_exp.so
for v in 60 6 7 8 9; do echo "S2D_PV_SMALL=$v"; S2D_PV_SMALL=$v b --workload c1 --steps 50; S2D_PV_SMALL=$v b --workload c4 --list-videos 32; done
for c in 5; do echo "S2D_PV_CTAS=$c"; S2D_PV_CTAS=$c b --steps 10; done
unset S2D_B200_LIB
python bench.py --videos 4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-k1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:point_votes_tab -s 1 -c 1 -f -o $out/r02_pv_v15 python bench.py --videos 4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-k1 > $out/ncu_pv15.log 2>&1
ls -la $out/r02_pv_v15.ncu-rep
