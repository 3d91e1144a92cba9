#!/bin/bash
# One-call A/B of library variants (tools/build_variant.sh) on the same box: config-2 / config-4 graph-replay step times, then a
# config-2 parity test + smoke on the fastest variant when it beats the in-tree build by more than 0.4 %.   usage: tools/ab_variants.sh TAG v1 v2 ...
tag=$1; shift
o=gpurun_out/${tag}_ab.log
: > $o
run() {
  if [ -z "$1" ]; then env -u MSTCN_B200_LIB python tools/quick_step.py --config $2 --steps 60 --reps 4 2>&1 | tail -1 | tee -a $o
  else MSTCN_B200_LIB=$1 python tools/quick_step.py --config $2 --steps 60 --reps 4 2>&1 | tail -1 | tee -a $o; fi
}
run "" 2
for v in "$@"; do run variants/libmstcn_$v.so 2; done
run "" 2
for v in "$@"; do run variants/libmstcn_$v.so 2; done
run "" 4
last=""; for v in "$@"; do last=$v; done
[ -n "$last" ] && run variants/libmstcn_$last.so 4
best=$(python - "$o" <<'PY'
import re, sys
t = {}
for line in open(sys.argv[1]):
    m = re.match(r"lib=(\S+) config=2 ms/step=([\d.]+)", line)
    if m:
        t.setdefault(m.group(1), []).append(float(m.group(2)))
base = min(t.get("default", [1e9]))
cand = {k: min(v) for k, v in t.items() if k != "default"}
k = min(cand, key=cand.get) if cand else ""
print(k if cand and cand[k] < base * 0.996 else "")
PY
)
echo "best variant: '${best}'" | tee -a $o
if [ -n "$best" ]; then
  # a quick parity check of the winner (the full suite runs on the final in-tree build in the next call)
  MSTCN_B200_LIB=$best timeout 120 python -m pytest tests/test_gpu_full_size.py -x -q -k "config2_train_mode" > gpurun_out/${tag}_pytest_variant.log 2>&1
  echo "pytest($best, config-2 parity) rc=$?" | tee -a $o
  tail -2 gpurun_out/${tag}_pytest_variant.log
  MSTCN_B200_LIB=$best python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a $o
else
  python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a $o
fi
