set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_shipped_configs.py -x -q -k "gvp" > gpurun_out/r02b_pytest_gvp.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02b_pytest_gvp.log
for v in "" _kcs0 _siluplain; do
  echo "=== variant '$v'"
  KPD_LIB=keypoint_diffusion_b200/libkpdiff_b200$v.so KPD_GVP_SERIAL=1 timeout 200 python tools/tc_phase_times.py bf16x3 > gpurun_out/r02b_phase$v.txt 2>&1; echo "rc=$?"
  head -9 gpurun_out/r02b_phase$v.txt; grep -A6 "edge CTAs" gpurun_out/r02b_phase$v.txt
done
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ragged --no-shipped-ll-block --no-mode-blocks > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r02b_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
