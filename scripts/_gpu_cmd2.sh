mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r01s_bench2.json 2> gpurun_out/r01s_bench2.err; echo "bench2 rc=$?"; tail -c 1500 gpurun_out/r01s_bench2.json; tail -5 gpurun_out/r01s_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r01s_ref2.json 2> gpurun_out/r01s_ref2.err; echo "ref2 rc=$?"; tail -c 1200 gpurun_out/r01s_ref2.json
