cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --workload egnn_20kp > gpurun_out/r02_final_bench_egnn_20kp.json 2> gpurun_out/err1.txt; echo "rc=$?"
timeout 900 python bench.py --workload egnn_20kp_c1 > gpurun_out/r02_final_bench_egnn_20kp_c1.json 2> gpurun_out/err2.txt; echo "rc=$?"
timeout 1500 python bench.py --workload egnn_all_atom --steps 1 --warmup 1 --no-mode-blocks --no-ragged --no-shipped-ll-block > gpurun_out/r02_final_bench_egnn_all_atom.json 2> gpurun_out/err3.txt; echo "rc=$?"
for n in 1 16 256 1024 4096; do
timeout 900 python bench.py --workload gvp_ca --ligands $n --steps 1 --warmup 1 --no-mode-blocks --no-cpu-baseline --no-roofline --no-ragged --no-shipped-ll-block > gpurun_out/r02_final_bench_gvp_ca_$n.json 2> gpurun_out/err4.txt; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],2), round(d['e2e']['value'],2), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
