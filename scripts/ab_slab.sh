#!/bin/bash
# Two-GPU check of the slab path on a GPU box: slab parity tests (in-process slabs and NCCL), then scripts/bench_slab.py
# at 257^3 with the halo exchange overlapped (HH_HALO_OVERLAP=1) and not.  Usage: gpurun --gpus 2 -- 'bash scripts/ab_slab.sh'
timeout 200 python -m pytest tests/test_gpu_slab.py tests/test_gpu_slab_nccl.py -q --timeout 150 2>&1 | tail -8
for p in 1 0; do
HH_HALO_OVERLAP=$p timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/bench_slab.py --grid 257 --nrhs 8 --steps 2 --warmup 1 2>gpurun_out/ab_err_$p.log | grep metric > gpurun_out/slab_n2_257_overlap$p.json
python - <<PY
import json
d=json.load(open("gpurun_out/slab_n2_257_overlap$p.json"))
print("overlap=$p", d["ms_per_step"], d["config"]["iterations"], d["config"]["true_relres_max"], {k:(v["launches"],v["share"],v["avg_ms"]) for k,v in d["per_kernel_rank0"].items() if k in ("halo_exchange","allreduce","fine_jacobi","coarse_jacobi","coarse_apply")})
PY
tail -2 gpurun_out/ab_err_$p.log
done
