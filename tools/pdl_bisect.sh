#!/bin/bash
# Localise the cross-kernel race of the flag-linked launches (MSTCN_PDL=1): alternating-batch graph replays at B=64 under
# every subset of kernel-to-kernel links.  Usage (GPU box): bash tools/pdl_bisect.sh [iters] > gpurun_out/pdl_bisect.log
IT=${1:-1200}
run() { echo "== $*"; env "$@" python tools/graph_stress.py --videos 64 --iters $IT 2>&1 | tail -4; }
run MSTCN_PDL=1
run MSTCN_PDL=1 MSTCN_DF=0
run MSTCN_PDL=1 MSTCN_DF_FWD=0
run MSTCN_PDL=1 MSTCN_DF_OFF=15
run MSTCN_PDL=1 MSTCN_DF_FWD=0 MSTCN_DF_OFF=14
run MSTCN_PDL=1 MSTCN_DF_FWD=0 MSTCN_DF_OFF=13
run MSTCN_PDL=1 MSTCN_DF_FWD=0 MSTCN_DF_OFF=11
run MSTCN_PDL=1 MSTCN_DF_FWD=0 MSTCN_DF_OFF=7
