#!/bin/bash
# Is the SASS of one kernel in the built library identical to what a given commit's point_votes.cu compiles to?
# profiles/k2_traffic.json (DRAM bytes of the C2 votes launch, ncu) was captured at commit 7d5f9d4; bench.py reports it while
# s2d_version() is unchanged, and the version is only kept when this check says SAME for the benchmarked kernel:
#   tools/sass_same.sh 7d5f9d4 '_ZN3s2d22point_votes_tab_kernelILi128ELi32ELi6ELb0ELi2EEEvPK14s2d_video_descPK4int4iPiS7_S7_PKh'
set -e
trap 'rm -rf "$tmp"' EXIT
commit=$1; fn=$2; root=$(cd "$(dirname "$0")/.." && pwd); tmp=$(mktemp -d)
mkdir -p $tmp/inc
for f in point_votes.cu common.cuh dbscan.cuh; do git -C $root show $commit:s2d_b200/csrc/$f > $tmp/$f; done
git -C $root show $commit:include/s2d_b200.h > $tmp/inc/s2d_b200.h
sed -i 's#../../include/s2d_b200.h#inc/s2d_b200.h#' $tmp/common.cuh $tmp/point_votes.cu
(cd $tmp && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I. -cubin -o old.cubin point_votes.cu)
strip() { grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##'; }
cuobjdump -sass -fun "$fn" $tmp/old.cubin | strip > $tmp/old.sass
cuobjdump -sass -fun "$fn" $root/s2d_b200/libs2d_b200.so 2>/dev/null | strip > $tmp/new.sass
test -s $tmp/old.sass
if cmp -s $tmp/old.sass $tmp/new.sass; then echo "SAME ($(wc -l < $tmp/new.sass) instructions)"; else echo DIFFERENT; exit 1; fi
