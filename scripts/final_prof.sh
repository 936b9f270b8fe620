# Round-2 evidence for profiles/ (one GPU).  Every ncu command runs after the same command exited 0 without ncu.
set -x
O=gpurun_out/final2
mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --steps 20 --warmup 5 --no-extras > $O/bench_n1_steps20.json 2> $O/bench_n1_steps20.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
for w in ge_grbm_cd1_b20 mnist_rbm_cd1_b20 mnist_rbm_pcd1_b20; do python bench.py --workload $w --no-cpu-baseline --no-extras > $O/bench_$w.json 2>/dev/null; done
for w in rbm_784x500_b8192_pcd1_tf32 rbm_784x500_b8192_pcd10_tf32; do python bench.py --workload $w --no-cpu-baseline --no-extras --steps 40 --warmup 5 > $O/bench_$w.json 2>/dev/null; done
python scripts/skinny_perf.py ge_b10_pcd1 ge_b10_cd1 ge_b20_cd1 ge_b10_pcd5 mnist_b20_cd1 dbn1000_b20_cd1 sm_b20_cd1 > $O/skinny_perf.txt 2>&1
python scripts/mid_batch.py > $O/mid_batch.txt 2>&1
(MDBN_SKINNY_TIMING=1 python scripts/skinny_perf.py ge_b10_pcd1 2>&1 | grep timeline | tail -2) > $O/timeline.txt
timeout 200 python scripts/small_layers.py > $O/small_layers.txt 2>&1
# launch list of the default bench command
python bench.py --steps 34 --warmup 5 --no-cpu-baseline --no-extras > $O/b34.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 34 --warmup 5 --no-cpu-baseline --no-extras > $O/ncu_l.log 2>&1
# the chained kernel: full set, and DRAM traffic with the caches left alone (what a step really moves)
python scripts/ncu_chain.py > $O/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:cd_skinny -s 2 -c 1 -f -o $O/chain python scripts/ncu_chain.py > $O/ncu_f.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --cache-control none --clock-control none -k regex:cd_skinny -s 1 -c 3 --csv --log-file $O/traffic.csv python scripts/ncu_chain.py > $O/ncu_t.log 2>&1
tail -2 $O/ncu_f.log
