# Multi-GPU lines of profiles/ (run under `gpurun --gpus N`, N = 2, 4 or 8): the driver-style bench line (weak scaling of the
# default workload + mdbn_aml_wallclock_s + the data-parallel legs) and the config-5 sweep sharded over the N ranks.
set -x
N=${1:-2}
O=gpurun_out/mg
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --steps 204 --warmup 34 --no-cpu-baseline 2> $O/bench_n$N.err | tail -1 > $O/bench_n$N.json
if [ "$2" = "sweep" ]; then
  $TR --nproc-per-node $N --master-port $((29600+N)) scripts/config_sweeps.py 2> $O/sweep_n$N.err | tail -1 > $O/config_sweeps_n$N.json
fi
head -c 600 $O/bench_n$N.json
