# Round-end evidence for profiles/ (one GPU).  Every ncu command runs after the same command exited 0 without ncu.
set -x
mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err
for w in ge_grbm_cd1_b20 mnist_rbm_cd1_b20; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/final/bench_$w.json 2>/dev/null; done
(timeout 60 python scripts/skinny_timeline.py 2>&1 | grep timeline | sed -n "4p;8p;12p"; echo COLD; COLD=1 timeout 60 python scripts/skinny_timeline.py 2>&1 | grep timeline | sed -n "16p;32p;48p") > gpurun_out/final/timeline.txt
(timeout 100 python scripts/skinny_timeline_small.py 2>&1 | grep timeline | sed -n "4p;8p;12p;16p") > gpurun_out/final/timeline_small.txt
timeout 200 python scripts/small_layers.py > gpurun_out/final/small_layers.txt 2>&1
python scripts/mdbn_aml_wallclock.py 2>/dev/null | tail -1 > gpurun_out/final/mdbn_n1.json
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/final/b30.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/final/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cd_skinny --launch-skip 8 -c 2 -o gpurun_out/final/skinny_final -f python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_f.log 2>&1
tail -2 gpurun_out/final/ncu_f.log
