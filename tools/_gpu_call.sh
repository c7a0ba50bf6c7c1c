cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --workload egnn_all_atom --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-mode-blocks --no-ragged --no-shipped-ll-block > gpurun_out/r02_final_bench_4gpu_egnn_all_atom.json 2> gpurun_out/err4g.txt; echo "rc=$?"
python -c "
import json
for l in open('gpurun_out/r02_final_bench_4gpu_egnn_all_atom.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['n_gpus'], d['ms_per_step'])
"
