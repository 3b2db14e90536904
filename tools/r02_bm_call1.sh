#!/bin/bash
# round 2, sparse-tile mode of the votes kernel (block-summary maps): correctness on both builds first, then the A/B of
# kernel configurations in one process (tools/r02_bm_ab.py), then one ncu capture of the kernel on the C1 video.
out=gpurun_out; mkdir -p $out
K='blockmap or block_maps or point_votes or ragged or golden or c1_shape or random_scenes or windowed or permutation or device_runner'
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_workloads.py -m gpu -x -q -k "$K" > $out/r02_bm_tests.log 2>&1; echo "exit $?" >> $out/r02_bm_tests.log; tail -3 $out/r02_bm_tests.log
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_check.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_workloads.py -m gpu -q -k "$K" > $out/r02_bm_tests_check.log 2>&1; echo "exit $?" >> $out/r02_bm_tests_check.log; tail -3 $out/r02_bm_tests_check.log
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_exp.so timeout 420 python tools/r02_bm_ab.py > $out/r02_bm_ab.json 2> $out/r02_bm_ab.err; echo "ab exit $?"; grep -c same_results $out/r02_bm_ab.json; cat $out/r02_bm_ab.err | tail -60
timeout 60 python bench.py --workload c1 --steps 50 --no-cpu --no-k1 > $out/r02_bench_c1_bm.json 2> $out/r02_bench_c1_bm.err; echo "c1 exit $?"; tail -c 600 $out/r02_bench_c1_bm.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:point_votes_tab -s 3 -c 1 -f -o $out/r02_pv_bm_c1 python bench.py --workload c1 --steps 2 --warmup 3 --no-e2e --no-cpu --no-k1 > $out/ncu_pv_bm.log 2>&1
ls -la $out/r02_pv_bm_c1.ncu-rep
