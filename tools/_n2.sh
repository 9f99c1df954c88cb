cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 tools/check_sharded.py 12 2>&1 | grep -v "^\*\*\*\|NCCL version\|OMP_NUM\|^$" | tail -3
for f in 512 1024; do
timeout 250 $TR --master-port 2955${f:0:1} bench.py --gpus 2 --streams 2 --frames-per-step $f > gpurun_out/gen_n2_f$f.json 2> gpurun_out/gen_n2_f$f.err; echo rc=$?
cut -c1-200 gpurun_out/gen_n2_f$f.json
done
CUDA_VISIBLE_DEVICES=0 timeout 250 python bench.py --frames-per-step 1024 > gpurun_out/bench_n1_f1024.json 2> gpurun_out/bench_n1_f1024.err; echo rc=$?
cut -c1-200 gpurun_out/bench_n1_f1024.json
