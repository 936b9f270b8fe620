set -x
O=gpurun_out/final4
mkdir -p $O
for w in ge_grbm_pcd1_b32 ge_grbm_cd1_b50 ge_grbm_pcd1_b100 mnist_rbm_cd1_b100; do python bench.py --workload $w --no-cpu-baseline --no-extras > $O/bench_$w.json 2>/dev/null; done
for w in rbm_784x500_b8192_pcd1_tf32 rbm_784x500_b8192_pcd10_tf32; do python bench.py --workload $w --no-cpu-baseline --no-extras --steps 40 --warmup 5 > $O/bench_$w.json 2>/dev/null; done
python scripts/mid_batch.py > $O/mid_batch.txt 2>&1
timeout 600 python scripts/config_sweeps.py > $O/config_sweeps_n1.json 2> $O/config_sweeps.err
python scripts/ncu_mid.py ge 128 0 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/launches_mid_ge_b128.csv python scripts/ncu_mid.py ge 128 0 > $O/ncu_m.log 2>&1
python scripts/ncu_mid.py ge 100 1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/launches_mid_ge_b100_pcd.csv python scripts/ncu_mid.py ge 100 1 > $O/ncu_m2.log 2>&1
python scripts/ncu_mid.py mnist 8192 1 1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $O/launches_b8192.csv python scripts/ncu_mid.py mnist 8192 1 1 > $O/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k tc_gemm_kernel -s 7 -c 1 -f -o $O/stats_split python scripts/ncu_mid.py ge 128 0 > $O/ncu_f.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k tc_gemm_kernel -s 9 -c 1 -f -o $O/up_plain_b8192 python scripts/ncu_mid.py mnist 8192 1 1 > $O/ncu_g.log 2>&1
tail -2 $O/ncu_f.log
