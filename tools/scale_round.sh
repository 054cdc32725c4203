#!/usr/bin/env bash
# Multi-GPU round: the 2-rank hardware check of the exchange, then bench.py at N GPUs with the in-kernel mailbox
# exchange and, for comparison, with the NCCL all-reduce. Usage: scale_round.sh TAG N
tag="${1:-x}"; n="${2:-2}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${tag}_multi_pytest.log 2>&1
tail -2 gpurun_out/${tag}_multi_pytest.log
run() {  # $1 = label, env SBOD_PEER_EXCHANGE
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 300 --warmup 5 > gpurun_out/${tag}_bench_${n}gpu_$1.json 2> gpurun_out/${tag}_bench_${n}gpu_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_${n}gpu_$1.json").read().strip().splitlines()[-1])
print("$1", "n=$n value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "joined", round(d["details"]["ms_per_step_joined"],4),
      "train", round(d["details"]["ms_train_half"],4), "eval", round(d["details"]["ms_eval_half"],4), "e2e", round(d["e2e"]["value"]), d["e2e"]["h2d_gbs_per_rank_all_ranks_copying"])
PY
}
run mailbox
SBOD_PEER_EXCHANGE=0 run nccl
