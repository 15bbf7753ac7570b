#!/bin/bash
# Short GPU-box visit: compute-sanitizer memcheck over the squaring-step kernel variants that rely on
# address redirects / lane hand-offs (parity tests cannot see an out-of-bounds READ), then the other
# single-GPU workloads on the current default.  Usage (through gpurun): bash scripts/gpu_exp_memcheck.sh <tag>
TAG=${1:-r01o}
O=gpurun_out
mkdir -p $O
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 \
    python -m pytest tests/test_gpu_kernels.py -q -x -k "test_morph_field and (776 or 256-)" > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/${TAG}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/${TAG}_memcheck.log | tail -5
for wl in c3 c2; do
  timeout 70 python bench.py --workload $wl --no-cpu-baseline --steps 50 > $O/${TAG}_bench_${wl}.json 2>> $O/${TAG}_bench.err; echo "bench $wl rc=$?"
done
python - <<PY
import json
for n in ("c3", "c2"):
    try:
        j = json.loads(open("$O/${TAG}_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value", j["value"], "ms/step", j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], j["roofline"]["kernel"])
    except Exception as e:
        print(n, "unreadable:", e)
PY
