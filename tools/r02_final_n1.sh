#!/bin/bash
# round 2: final single-GPU evidence - full suites on both builds, bench lines, ncu captures
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r02_gputest_product.log 2>&1; echo "exit $?" >> $out/r02_gputest_product.log; tail -3 $out/r02_gputest_product.log
S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_check.so python -m pytest tests -m gpu -q > $out/r02_gputest_boundscheck.log 2>&1; echo "exit $?" >> $out/r02_gputest_boundscheck.log; tail -3 $out/r02_gputest_boundscheck.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 > $out/r02_bench_c2_final.json 2> $out/r02_bench_c2_final.err; tail -c 400 $out/r02_bench_c2_final.err
python bench.py --workload target --no-cpu > $out/r02_bench_target_480p.json 2>> $out/r02_bench_c2_final.err
python bench.py --workload c1 --steps 50 > $out/r02_bench_c1.json 2>> $out/r02_bench_c2_final.err
python bench.py --workload c1 --steps 50 --graph --no-cpu > $out/r02_bench_c1_cuda_graph.json 2>> $out/r02_bench_c2_final.err
python bench.py --impl reference --steps 1 --warmup 0 > $out/r02_bench_c2_reference_arm.json 2>> $out/r02_bench_c2_final.err
for f in r02_bench_c2_final r02_bench_target_480p r02_bench_c1 r02_bench_c1_cuda_graph r02_bench_c2_reference_arm; do python - $out/$f.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
r=d.get('roofline') or {}
print(sys.argv[1].split('/')[-1], round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'k2', round(r.get('frac',0),4), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('h2d_probe_gbs_per_gpu'), d.get('stage_ms'), (d.get('overlap_gemm') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
PY
done
# launch list of one step (4 videos), then full captures of K2 on the whole C2 batch (DRAM traffic of the benchmarked launch),
# of the label histogram and of the Gram kernel
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r02_launches.csv -k regex:"label_hist|frame_tables|vis_reduce|binarize|db_|db1_|cluster_count|windows_kernel|pv_|point_votes|select_kernel|group_|video_status" -c 400 python bench.py --videos 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-k1 > $out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:point_votes_tab -s 3 -c 1 -f -o $out/r02_pv_v16_c2full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-k1 > $out/ncu_pv16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:label_hist -s 1 -c 1 -f -o $out/r02_label_hist python bench.py --videos 16 --steps 1 --warmup 3 --no-e2e --no-cpu --no-k1 > $out/ncu_lh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_labels2 -c 1 -f -o $out/r02_gram_v6 python tools/k1_one.py > $out/ncu_gram6.log 2>&1
ls -la $out/*.ncu-rep | tail -4
