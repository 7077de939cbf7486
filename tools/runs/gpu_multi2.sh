mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_cpp_api.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02z_bench2.json 2> gpurun_out/r02z_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02z_bench2.json').read().strip().splitlines()[-1])
print('N=2 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'reduce',d.get('reduce_check',{}).get('psnr_db'),'strong',round(d['strong_job']['wall_ms'],1))
for c in d['configs']: print('  ',c['workload'][:40],round(c['value'],1),round(c['e2e']['value'],1))
PY
