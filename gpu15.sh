mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_parity.py > gpurun_out/mgpu_parity_2.log 2>&1; echo "parity2 rc=$?"; tail -4 gpurun_out/mgpu_parity_2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/mgpu_parity.py > gpurun_out/mgpu_parity_8.log 2>&1; echo "parity8 rc=$?"; tail -4 gpurun_out/mgpu_parity_8.log
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 24 --warmup 6 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err;
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 24 --warmup 6 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  echo "n=$n rc=$?"; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/scale_n$n.json') if l.startswith('{')][-1]); print('n_gpus',d['n_gpus'],'img/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1))"
done
