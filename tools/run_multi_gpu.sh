#!/bin/bash
# 8-GPU / 2-GPU bench lines (weak scaling on the 1080p path, config 5 probes) + the 2-GPU equality test.
# usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/run_multi_gpu.sh'
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
for n in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 \
    > gpurun_out/final_bench_${n}gpu.json 2> gpurun_out/final_bench_${n}gpu.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload castle_512x256_probes \
  --steps 10 --warmup 3 > gpurun_out/final_bench_8gpu_probes.json 2> /dev/null
for f in final_bench_8gpu final_bench_2gpu final_bench_8gpu_probes; do
  python -c "
import json
d=json.load(open('gpurun_out/$f.json')); print('$f', d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['clocks'])"
done
