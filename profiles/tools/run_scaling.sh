#!/bin/bash
# usage (on the GPU box): bash profiles/tools/run_scaling.sh <tag>  -> gpurun_out/scale_<tag>_N.json for N = 1 2 4 8
tag=${1:-x}
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${tag}_1.json 2> gpurun_out/scale_${tag}_1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_${tag}_$n.json 2> gpurun_out/scale_${tag}_$n.err
done
python - <<'PY'
import json, glob, re, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "x"
PY
for n in 1 2 4 8; do grep -h '"metric"' gpurun_out/scale_${tag}_$n.json gpurun_out/scale_${tag}_$n.err 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('N=%d ms %.4f evals/s %.1f e2e_ms %.4f kernel_ms_host_call %.4f' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['extra']['kernel_ms_in_host_call']))
"; done
