#!/bin/bash
# round 2, one-warp-per-tile votes kernel (P <= 1024): correctness on the bounds-checked and the product build first, then the
# A/B against the label-table kernel in one process, the C1 bench line, and one ncu capture of the kernel (C4 slice).
out=gpurun_out; mkdir -p $out
K='point_votes or ragged or golden or c1_shape or random_scenes or windowed or permutation or device_runner or cuda_graph'
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_check.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_workloads.py -m gpu -x -q -k "$K" > $out/r02_warp_tests_check.log 2>&1; echo "exit $?" >> $out/r02_warp_tests_check.log; tail -4 $out/r02_warp_tests_check.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_workloads.py -m gpu -x -q -k "$K" > $out/r02_warp_tests.log 2>&1; echo "exit $?" >> $out/r02_warp_tests.log; tail -4 $out/r02_warp_tests.log
timeout 300 python tools/r02_k2_small_ab.py > $out/r02_k2_small_ab.json 2> $out/r02_k2_small_ab.err; echo "ab exit $?"; tail -20 $out/r02_k2_small_ab.err
timeout 60 python bench.py --workload c1 --steps 50 --no-cpu --no-k1 > $out/r02_bench_c1_warp.json 2> $out/r02_bench_c1_warp.err; echo "c1 exit $?"; tail -c 700 $out/r02_bench_c1_warp.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:point_votes_warp -s 1 -c 1 -f -o $out/r02_pv_warp_c4 python bench.py --workload c4 --list-videos 4 --steps 1 > $out/ncu_pv_warp.log 2>&1
ls -la $out/r02_pv_warp_c4.ncu-rep
