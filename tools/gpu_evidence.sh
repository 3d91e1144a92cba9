#!/bin/bash
# list of an eager step and full captures of the two top kernels.  Everything lands in gpurun_out/ (copy the summaries
# worth keeping into profiles/).   usage: tools/gpu_evidence.sh TAG
# (compute-sanitizer is closed on this GPU pool -- it answers with a refusal -- so there is no memcheck / racecheck leg;
#  tools/trapdbg.py and mstcn_debug_trap_report are the post-mortem for a wait that never completes.)
# tc_layer_kernel launches of tools/fwd_only.py: <0> <3> per stage -> the 9th is a forward chain; of tools/bwd_only.py: 8 forward
# + per stage <4> <2> <1> -> launch 30 is a backward chain of the second step.  tc_wgrad_kernel: 4 stage launches + the projection-mode
# launch per step -> launch 5 (skip 5) is the last stage's launch of the second step, launch 4 the first step's projection launch.
tag=${1:-r02}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/${tag}_pytest.log
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $o/${tag}_launches.csv python tools/bwd_only.py > $o/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_layer_kernel -s 8 -c 1 -f -o $o/${tag}_chain_fwd python tools/fwd_only.py > $o/${tag}_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 5 -c 1 -f -o $o/${tag}_wgrad python tools/bwd_only.py > $o/${tag}_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_layer_kernel -s 29 -c 1 -f -o $o/${tag}_chain_bwd python tools/bwd_only.py > $o/${tag}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
