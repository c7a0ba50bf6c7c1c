#!/bin/bash
# A/B sweep: concurrently sampled sub-batches x (dependency-graph convs | serial convs, GVP only); prints ligands/s
# usage: tools/sweep_sub_batches.sh [workload] [precision] ["0 1"]
WL=${1:-gvp_20kp}; PR=${2:-bf16x3}; SER=${3:-"0 1"}
for serial in $SER; do for n in 1 2 3 4 6; do
  v=$(KPD_GVP_SERIAL=$serial timeout 300 python bench.py --workload $WL --precision $PR --sub-batches $n --steps 2 --warmup 2 --no-cpu-baseline --no-roofline --no-mode-blocks 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['e2e']['value'],2), d['launches_per_reverse_step'])")
  echo "$WL $PR serial_convs=$serial sub_batches=$n ligands/s(value,e2e,launches/step): $v"
done; done
