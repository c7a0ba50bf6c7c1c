cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py > gpurun_out/r02_final2_bench_default.json 2> gpurun_out/r02_final2_bench_default.err; echo "bench rc=$?"
for n in 256 1024; do
timeout 900 python bench.py --workload gvp_ca --ligands $n --steps 1 --warmup 1 --no-mode-blocks --no-cpu-baseline --no-roofline --no-ragged --no-shipped-ll-block > gpurun_out/r02_final2_bench_gvp_ca_$n.json 2> gpurun_out/err4.txt; echo "rc=$?"
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_final2_bench_default.json'))
print('ours', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step'], d['cpu_baseline']['value'], d['ragged']['value'], d['shipped_ll_cutoff']['value'], d['cold_call'])
print({k:v['value'] for k,v in d['modes'].items()})
for n in (256,1024):
    r=json.load(open(f'gpurun_out/r02_final2_bench_gvp_ca_{n}.json')); print(n, r['value'], r['e2e']['value'], r['ms_per_step'])
PY
