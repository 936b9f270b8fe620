# Round-2 closing evidence (one GPU).  Every ncu command runs after the same command exited 0 without ncu.
set -x
O=gpurun_out/final5
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/tests.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --steps 20 --warmup 5 --no-extras > $O/bench_n1_steps20.json 2> $O/bench_n1_steps20.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
for w in ge_grbm_cd1_b20 mnist_rbm_cd1_b20 mnist_rbm_pcd1_b20; do python bench.py --workload $w --no-cpu-baseline --no-extras > $O/bench_$w.json 2>/dev/null; done
timeout 200 python scripts/small_layers.py > $O/small_layers.txt 2>&1
python scripts/mdbn_aml_wallclock.py > $O/mdbn_n1.json 2> $O/mdbn_n1.err
# launch lists + full captures of the two small-layer kernels
python scripts/ncu_small.py mnist && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_mid_mnist_b20.csv python scripts/ncu_small.py mnist > $O/ncu_l1.log 2>&1
python scripts/ncu_small.py me && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_tiny_me_k10.csv python scripts/ncu_small.py me > $O/ncu_l2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cd_mid -s 2 -c 1 -f -o $O/mid_mnist python scripts/ncu_small.py mnist > $O/ncu_f1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cd_tiny -s 2 -c 1 -f -o $O/tiny_me python scripts/ncu_small.py me > $O/ncu_f2.log 2>&1
cat $O/tests.log; tail -2 $O/ncu_f2.log
