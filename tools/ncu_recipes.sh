#!/bin/bash
# ncu recipes for this repo, to be run on the GPU box (gpurun -- 'bash tools/ncu_recipes.sh <recipe>').
# Rules learnt the hard way:
#   * never profile the full C2 bench without a kernel filter: ncu saves and restores all resident device memory
#     (60 GB of synthetic inputs) around every profiled launch, torch's input generators included - one such call
#     ran into a 15-minute limit. Use --videos 4 (or tools/k1_one.py) and -k regex:<kernel> -c <n>.
#   * run the same command once WITHOUT ncu first; a number printed under ncu is never a bench value.
set -euo pipefail
out=gpurun_out
mkdir -p "$out"
case "${1:-help}" in
  launches)   # launch list of our kernels over two steps of a 4-video batch
    python bench.py --videos 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-k1 > /dev/null
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches.csv" \
        -k regex:'label_hist|frame_tables|vis_reduce|binarize|db_|cluster_count|windows_kernel|pv_|point_votes|select_kernel|group_|video_status' \
        -c 400 python bench.py --videos 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-k1 > "$out/ncu_launches.log" 2>&1
    ;;
  votes)      # full capture of the votes kernel (K2) on a 4-video batch
    python bench.py --videos 4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-k1 > /dev/null
    ncu --set full --clock-control none --import-source on -k regex:point_votes_tab -s 1 -c 1 -f -o "$out/pv" \
        python bench.py --videos 4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-k1 > "$out/ncu_pv.log" 2>&1
    ;;
  dbscan)     # the three DBSCAN passes of the group stage (launches 4-6 of a step are stage D's)
    python bench.py --videos 4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-k1 > /dev/null
    ncu --set full --clock-control none --import-source on -k regex:db_pass_kernel -s 3 -c 3 -f -o "$out/db_group" \
        python bench.py --videos 4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-k1 > "$out/ncu_db.log" 2>&1
    ;;
  gram)       # the Gram kernel (K1) on one C2 video
    python tools/k1_one.py > /dev/null
    ncu --set full --clock-control none --import-source on -k regex:gram_labels2 -c 1 -f -o "$out/gram" \
        python tools/k1_one.py > "$out/ncu_gram.log" 2>&1
    ;;
  *)
    echo "usage: $0 launches|votes|dbscan|gram   (summaries: python tools/ncu_summary.py <rep> <tiles>; tools/ncu_bylines.py <source csv> <tiles>)"
    ;;
esac
