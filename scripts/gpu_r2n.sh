#!/bin/bash
TAG=$1
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "anatomy or final_pass or intensity or morph_field" 2>&1 | tail -4
timeout 500 python bench.py --steps 100 --warmup 5 --profile-out $O/${TAG}_bench_profile.json > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "advk_ms_per_step")}, d["e2e"], d["roofline"]["frac"], d["cuda_eager_baseline"], d["cpu_baseline"]["value"])
print(d["roofline_scopes"])
PY
