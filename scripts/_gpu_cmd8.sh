mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
for n in 4 8; do
  if [ $n -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/r01s_bench$n.json 2> gpurun_out/r01s_bench$n.err; echo "bench$n rc=$?"
    python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    l=json.loads(open('gpurun_out/r01s_bench%s.json'%n).read().strip().splitlines()[-1])
    print({k:l[k] for k in ['value','n_gpus','ms_per_step','e2e','clocks']})
except Exception as e:
    print("no json", e); print(open('gpurun_out/r01s_bench%s.err'%n).read()[-1500:])
PY
  fi
done
