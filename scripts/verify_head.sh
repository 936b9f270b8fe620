# Round-end check of the committed state (one GPU): GPU parity suite, smoke(), the default bench line, the driver-sized one
set -x
O=gpurun_out/verify
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --steps 20 --warmup 5 --no-extras > $O/bench_steps20.json 2> $O/bench_steps20.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
for w in mnist_rbm_cd1_b20 mnist_rbm_pcd1_b20; do python bench.py --workload $w --no-cpu-baseline --no-extras > $O/bench_$w.json 2>/dev/null; done
cat $O/tests.log; tail -2 $O/smoke.log; cut -c1-300 $O/bench_n1.json
