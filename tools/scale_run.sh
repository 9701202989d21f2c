#!/bin/bash
# Scaling runs of bench.py on one box: C1 at 1/2/4/8 GPUs, C3 at 2/4/8 GPUs. Writes gpurun_out/scale_*.json.
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], sys.argv[3], '%.3e/s' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'walk %.3f build %.3f' % (d['roofline']['walk_ms'], d['roofline']['build_ms']))" "$@"; }
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_c1_1.json; show gpurun_out/scale_c1_1.json C1 1
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_c1_$n.json; show gpurun_out/scale_c1_$n.json C1 $n
done
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload gas16m --steps 5 --warmup 2 2>/dev/null | tail -1 > gpurun_out/scale_c3_$n.json; show gpurun_out/scale_c3_$n.json C3 $n
done
