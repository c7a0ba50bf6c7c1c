cd $GRAFT_REPO_ROOT
timeout 400 python bench.py --workload gvp_ca --ligands 16 --steps 2 --warmup 2 --no-cpu-baseline --no-ragged --no-shipped-ll-block --no-mode-blocks --no-roofline > gpurun_out/r02l_bench_ca16.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r02l_bench_ca16.json')); print('gvp_ca 16', d['value'], d['e2e']['value'])"
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ragged --no-shipped-ll-block --no-mode-blocks --no-roofline > gpurun_out/r02l_bench_gvp.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r02l_bench_gvp.json')); print('gvp', d['value'], d['e2e']['value'])"
timeout 600 python -m pytest tests -x -q -m gpu -k "gvp or sample or loop" > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02l_pytest.log
