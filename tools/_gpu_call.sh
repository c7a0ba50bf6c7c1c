cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02g_pytest.log
for sb in 1 2 3 4; do
timeout 400 python bench.py --workload egnn_20kp --steps 2 --warmup 3 --sub-batches $sb --no-cpu-baseline --no-ragged --no-shipped-ll-block --no-mode-blocks --no-roofline > gpurun_out/r02g_bench_egnn_sb$sb.json 2> gpurun_out/r02g_bench_egnn.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r02g_bench_egnn_sb$sb.json')); print('egnn sub-batches $sb', d['value'], d['e2e']['value'])"
done
