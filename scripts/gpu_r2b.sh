#!/bin/bash
# Visit: parity suite, smoke, the default bench run (all legs) and the reference arm.  Usage: bash scripts/gpu_r2b.sh <tag> [noref]
TAG=$1; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log; tail -2 $O/${TAG}_smoke.log
( time timeout 900 python bench.py --profile-out $O/${TAG}_bench_profile.json > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "advk_ms_per_step")}, d["e2e"]["value"])
print(json.dumps(d["roofline"], indent=1))
print(d["cuda_eager_baseline"]); print(d["cpu_baseline"])
PY
[ "$2" = "noref" ] || { ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err ) 2>&1 | grep real; cat $O/${TAG}_bench_ref.json | cut -c1-600; }
