#!/bin/bash
export S2D_B200_LIB=$PWD/s2d_b200/libs2d_b200_exp.so
for c in 1.0 0.85 0.7 0.55 0.4; do echo "diag cost $c"; S2D_GRAM_DIAG_COST=$c python tools/k1_bench.py 2>/dev/null | python -c "
import json,sys; a=json.load(sys.stdin)
print({k:(round(v['ms_per_video'],4)) for k,v in a.items() if isinstance(v,dict)})"; done
