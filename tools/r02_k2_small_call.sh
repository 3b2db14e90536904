#!/bin/bash
# round 2, one-warp-per-tile votes kernel (P <= 1024): correctness on the bounds-checked build first (the product build runs
# the whole suite in tools/r02_final2_n1.sh), then the A/B against the label-table kernel in one process and one ncu capture
# of the kernel on a 4-video slice of the C4 mixture.
out=gpurun_out; mkdir -p $out
K='point_votes or ragged or golden or c1_shape or random_scenes or windowed or permutation or device_runner or cuda_graph'
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_check.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_workloads.py -m gpu -x -q -k "$K" > $out/r02_warp2_tests_check.log 2>&1; echo "exit $?" >> $out/r02_warp2_tests_check.log; tail -4 $out/r02_warp2_tests_check.log
timeout 300 python tools/r02_k2_small_ab.py > $out/r02_k2_small_ab2.json 2> $out/r02_k2_small_ab2.err; echo "ab exit $?"; tail -20 $out/r02_k2_small_ab2.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:point_votes_warp -s 1 -c 1 -f -o $out/r02_pv_warp2_c4 python bench.py --workload c4 --list-videos 4 --steps 1 > $out/ncu_pv_warp2.log 2>&1
ls -la $out/r02_pv_warp2_c4.ncu-rep
