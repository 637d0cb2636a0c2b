#!/bin/bash
# bench.py at N = 1, 2, 4, 8 back to back on one 8-GPU box (gpurun --gpus 8): gpurun_out/scale_n<N>.json
# optional: ACM_LM_TRACE run of the 8-GPU solve afterwards (scale_lm_trace_n8.log)
python bench.py --gpus 1 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
port=29520
for n in 2 4 8; do
  port=$((port + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
done
for n in 1 2 4 8; do tail -c 300 gpurun_out/scale_n$n.json | head -c 0; python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
lm = d["lm_conversion"]
print($n, round(d["value"] / 1e9, 1), "Gpts/s", d["ms_per_step"], "ms/step | LM", lm["ms"], "ms device", lm["device_ms"], "| checks", d["check"].get("ok"), lm.get("vs_oracle", {}).get("ok"), lm.get("rank_identical"), "| undistort", d["undistort_batch"].get("frames_per_s"))
PY
done
