cd /root/repo
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 tools/check_sharded.py 12 2>&1 | grep -v "^\*\*\*\|NCCL version\|OMP_NUM\|^$" | tail -3
timeout 300 $TR --master-port 29552 bench.py --gpus $N > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo rc=$?
cut -c1-220 gpurun_out/final_bench_n$N.json
