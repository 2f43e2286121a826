N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err; echo "bench N=$N rc=$?"
tail -2 gpurun_out/r2c_bench_n$N.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/r2c_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d.get('sharded_equals_single'))"
