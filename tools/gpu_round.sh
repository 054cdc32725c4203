#!/usr/bin/env bash
# One GPU-box round: parity tests, per-kernel timings of the BASELINE shapes, the bench line. Usage: gpu_round.sh TAG [quick]
tag="${1:-x}"
mode="${2:-full}"
mkdir -p gpurun_out
if [ "$mode" != "quick" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  tail -3 gpurun_out/${tag}_pytest.log
fi
{
  KB_PREFILL=1 timeout 300 python tools/kbench.py ssd512_canonical 32 81 100
  if [ "$mode" != "quick" ]; then
    KB_CLS=3 KB_REG=3 timeout 300 python tools/kbench.py retinanet640 32 81 100
    KB_PREFILL=1 timeout 300 python tools/kbench.py refinedet512 32 4 200
    KB_CLS=2 KB_REG=0 timeout 300 python tools/kbench.py ssd300 32 21 20
  fi
} > gpurun_out/${tag}_kb.log 2>&1
cat gpurun_out/${tag}_kb.log | grep -v "^==" | tail -40
timeout 600 python bench.py > gpurun_out/${tag}_bench2.json 2> gpurun_out/${tag}_bench2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench2.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", d["ms_per_step"], {k:round(v,4) for k,v in d["details"].items() if k.startswith("ms_")})
print({k:round(v.get("ms",0),4) for k,v in d["roofline"]["kernels"].items()})
PY
