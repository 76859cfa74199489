#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/t_mg.log 2>&1; echo "multigpu rc=$?"; grep -E "DDP|2-GPU|passed|failed|skipped|Error|error" gpurun_out/t_mg.log | tail -12
for n in 1 2; do
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2>gpurun_out/scale.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2>gpurun_out/scale.err; fi
  echo "scale n=$n rc=$?"; grep '^{' gpurun_out/scale_$n.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), d['clocks'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --workload train --steps 2 --warmup 2 > gpurun_out/train_2.json 2>gpurun_out/scale.err; echo "train 2gpu rc=$?"; grep '^{' gpurun_out/train_2.json | tail -1 | cut -c1-400
