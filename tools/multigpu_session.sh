#!/bin/bash
# One 8-GPU lease: slab-vs-replicated-vs-oracle parity at 8 and 4 ranks (tools/dist_check.py), then the scaling
# bench lines.  Logs land in gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
LOG=gpurun_out/r2_dist_n8_n4.log
: > $LOG
run() {  # nproc shape... -- gl agg [--oracle]
  np=$1; shift
  echo "=== nproc $np: $*" | tee -a $LOG
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 200)) tools/dist_check.py "$@" 2>&1 | grep -v "^W1\|^\[W\|Warning\|warn" >> $LOG
  echo "rc=${PIPESTATUS[0]}" | tee -a $LOG
}
run 8 --shape 64 64 64 --gl 3 --agg 4096 --oracle
run 8 --shape 1024 16 1024 --gl 4 --agg 65536
run 8 --shape 256 256 --gl 4 --agg 1024 --oracle
run 4 --shape 1024 16 1024 --gl 4 --agg 65536
grep -c "FAIL" $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29908 \
    bench.py --gpus 8 --steps 10 --warmup 3 --reps 3 > gpurun_out/r2_bench_n8_pull.json 2> gpurun_out/r2_bench_n8_pull.err
echo "bench n=8 pull rc=$?"; tail -c 300 gpurun_out/r2_bench_n8_pull.json
OMG_AGGLOMERATE_BELOW=2097152 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29950 \
    bench.py --gpus 8 --steps 10 --warmup 3 --reps 2 --no-extra > gpurun_out/r2_bench_n8_pull_agg2m.json 2> gpurun_out/r2_bench_n8_pull_agg2m.err
echo "bench n=8 pull agg2m rc=$?"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
