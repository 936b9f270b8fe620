# Round-2 (second half) evidence for profiles/ (one GPU).  Every ncu command runs after the same command exited 0 without ncu.
set -x
O=gpurun_out/final3
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/tests.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --steps 20 --warmup 5 --no-extras > $O/bench_n1_steps20.json 2> $O/bench_n1_steps20.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
for w in ge_grbm_cd1_b20 mnist_rbm_cd1_b20 mnist_rbm_pcd1_b20 ge_grbm_pcd1_b32 ge_grbm_cd1_b50 ge_grbm_pcd1_b100 mnist_rbm_cd1_b100; do python bench.py --workload $w --no-cpu-baseline --no-extras > $O/bench_$w.json 2>/dev/null; done
for w in rbm_784x500_b8192_pcd1_tf32 rbm_784x500_b8192_pcd10_tf32; do python bench.py --workload $w --no-cpu-baseline --no-extras --steps 40 --warmup 5 > $O/bench_$w.json 2>/dev/null; done
python scripts/skinny_perf.py ge_b10_pcd1 ge_b10_cd1 ge_b20_cd1 ge_b10_pcd5 mnist_b20_cd1 dbn1000_b20_cd1 sm_b20_cd1 > $O/skinny_perf.txt 2>&1
python scripts/mid_batch.py > $O/mid_batch.txt 2>&1
timeout 200 python scripts/small_layers.py > $O/small_layers.txt 2>&1
timeout 600 python scripts/config_sweeps.py > $O/config_sweeps_n1.json 2> $O/config_sweeps.err
# launch lists (warm caches): the default bench command, a mid-batch step, a large-batch step
python bench.py --steps 34 --warmup 5 --no-cpu-baseline --no-extras > $O/b34.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_skinny_bench.csv python bench.py --steps 34 --warmup 5 --no-cpu-baseline --no-extras > $O/ncu_l.log 2>&1
python scripts/ncu_mid.py ge 128 0 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/launches_mid_ge_b128.csv python scripts/ncu_mid.py ge 128 0 > $O/ncu_m.log 2>&1
python scripts/ncu_mid.py ge 100 1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/launches_mid_ge_b100_pcd.csv python scripts/ncu_mid.py ge 100 1 > $O/ncu_m2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/launches_b8192.csv python bench.py --workload rbm_784x500_b8192_pcd1_tf32 --no-cpu-baseline --no-extras --steps 4 --warmup 3 > $O/ncu_b.log 2>&1
# the statistics GEMM with the fused update (split-TF32, persistent): full set
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k tc_gemm_kernel -s 7 -c 1 -f -o $O/stats_split python scripts/ncu_mid.py ge 128 0 > $O/ncu_f.log 2>&1
# the chained skinny kernel: full set (unchanged kernel; refreshed for the record)
python scripts/ncu_chain.py > $O/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:cd_skinny -s 2 -c 1 -f -o $O/chain python scripts/ncu_chain.py > $O/ncu_c.log 2>&1
tail -2 $O/ncu_f.log
