#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N tag
NG=${1:-2}; TAG=${2:-r02v}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_${NG}gpu_$TAG.txt; nproc >> gpurun_out/topo_${NG}gpu_$TAG.txt; free -g | head -2 >> gpurun_out/topo_${NG}gpu_$TAG.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_c4_${NG}gpu_$TAG.json 2> gpurun_out/bench_c4_${NG}gpu_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_${NG}gpu_$TAG.json')); print({k: d[k] for k in ('value','ms_per_step','scaling','n_gpus')}); print(d['roofline']['frac'], d['roofline']['kernel_ms']); e=d['e2e']; print({k: e.get(k) for k in ('value','ms_per_step','symbols_per_gpu','d2h_gbs_achieved','pcie_probe','d2h_frac_of_link')})"
tail -2 gpurun_out/bench_c4_${NG}gpu_$TAG.err
timeout 600 python -m pytest tests -m gpu -q -k "multi or shard" 2>&1 | tail -3
timeout 300 python scripts/probe_pcie.py 2>&1 | tee gpurun_out/probe_pcie_${NG}gpu_$TAG.json | cut -c1-600
