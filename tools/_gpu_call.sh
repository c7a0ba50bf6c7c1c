cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02f_pytest.log
KPD_LIB=keypoint_diffusion_b200/libkpdiff_b200_tcgt.so timeout 120 python tools/tc_linear_bench.py 2>&1 | tail -7
timeout 400 python bench.py --workload egnn_20kp --steps 2 --warmup 3 --no-cpu-baseline --no-ragged --no-shipped-ll-block --no-mode-blocks > gpurun_out/r02f_bench_egnn.json 2> gpurun_out/r02f_bench_egnn.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r02f_bench_egnn.json')); print('egnn', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline'].get('kernel_share_of_step'))"
