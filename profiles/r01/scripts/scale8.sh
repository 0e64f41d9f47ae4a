cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; echo "rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_${N}gpu.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench_${N}gpu.log').read()[-1500:])
else:
    d=json.loads(l[-1]); print('n_gpus %d: value %.1f Gpts/s ms/step %.2f e2e %.1f Gpts/s' % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9), d['clocks'])
PY
