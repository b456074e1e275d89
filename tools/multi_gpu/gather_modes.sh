# usage: NGPU=8 bash tools/multi_gpu/gather_modes.sh   (run under gpurun --gpus N): step time with no gather, the peer-write gather and the NCCL gather
for g in none peer nccl; do
  if [ $g = none ]; then G=""; else G="--gather $g"; fi
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-8} --steps 2000 --warmup 200 --no-cpu $G 2>gpurun_out/gather_$g.err > gpurun_out/gather_$g.out
  echo "mode $g rc=$?"
  grep -o '"value": [0-9.e+]*, "unit": "env-steps/s", "n_gpus": [0-9]*[^}]*"ms_per_step": [0-9.]*' gpurun_out/gather_$g.out | head -1
  grep -m2 "rror" gpurun_out/gather_$g.err | cut -c1-200
done
