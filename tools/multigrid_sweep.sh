#!/bin/bash
# Throughput of the training step on the multigrid clip shapes of BASELINE config 3: per GPU
# (B, T, H) with BN splits = 2 x long-cycle scale.  One JSON line per shape into gpurun_out/<round>_multigrid_shapes.jsonl
R=${1:-r01}
OUT=gpurun_out/${R}_multigrid_shapes.jsonl
: > $OUT
for cfg in "256 4 111 16" "128 4 158 16" "128 8 111 8" "64 8 158 8" "128 8 112 4" "64 8 158 4" "32 8 224 4" "64 16 112 2" "32 16 158 2" "16 16 224 2"; do
  set -- $cfg
  timeout 300 python bench.py --batch $1 --frames $2 --crop $3 --bn-splits $4 --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $OUT || echo "{\"failed\": \"$cfg\"}" >> $OUT
done
python - <<PY
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:100]); continue
    if "failed" in d: print(d); continue
    print(d["config"]["workload"][:90], "->", round(d["value"], 1), "clips/s", round(d["ms_per_step"], 2), "ms")
PY
