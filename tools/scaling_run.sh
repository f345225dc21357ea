#!/bin/bash
# usage: scaling_run.sh <tag> <gpus> [bench flags]: bench.py under torchrun on <gpus> GPUs of this box, JSON line to
# gpurun_out/<tag>_n<gpus>.json (default flags: --no-e2e --no-cpu-baseline)
tag=$1; n=$2; shift 2
flags=${*:---no-e2e --no-cpu-baseline}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 $flags > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${tag}_n$n.json").read().strip().splitlines()[-1])
    print($n, d["value"], (d.get("e2e") or {}).get("value"), d["counts_operator"]["tile_cells"], {k:(v["ms"],v["launches"]) for k,v in d["kernel_classes"].items()}, d["parity"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/${tag}_n$n.err").read()[-1500:])
EOF
