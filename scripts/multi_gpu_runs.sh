# Multi-GPU lines of profiles/ (run under `gpurun --gpus 4`)
set -x
mkdir -p gpurun_out/mg
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4; do
  $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/mg/bench_n$n.json
  $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --workload rbm_784x500_b8192_pcd1_tf32 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/mg/dp_n$n.json
done
python bench.py --workload rbm_784x500_b8192_pcd1_tf32 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/mg/dp_n1.json
python scripts/mdbn_aml_wallclock.py 2>/dev/null | tail -1 > gpurun_out/mg/mdbn_n1.json
$TR --nproc-per-node 3 --master-port 29710 scripts/mdbn_aml_wallclock.py 2>/dev/null | tail -1 > gpurun_out/mg/mdbn_n3.json
head -c 300 gpurun_out/mg/*.json
