#!/bin/bash
# GPU call 3 (2 GPUs): NCCL slab path with the one-all-reduce-per-step level GMRES; bench line at N=2 with 32 RHS per step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02b
mkdir -p $O
exec > $O/call3.log 2>&1
set -x
date
nvidia-smi -L
free -g | head -2
timeout 400 python -m pytest tests/test_gpu_slab_nccl.py -m gpu -x -q 2>&1 | tail -8
date
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 --e2e-steps 1 > $O/bench3_n2.json 2> $O/bench3_n2.err
date
tail -n 5 $O/bench3_n2.err
python - $O/bench3_n2.json <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "n_gpus")})
        print("e2e", d.get("e2e"))
        print("ps", d.get("e2e_point_sources"))
        print("slab", d.get("slab"))
        print(d["config"])
PY
date
