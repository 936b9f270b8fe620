# mid kernel: parity tests, phase timeline, us per step against the row-slab kernel on the same layers
set -x
O=gpurun_out/mid8
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "mid" 2>&1 | tail -5 > $O/tests.log
cat $O/tests.log
python scripts/mid_timeline.py 2>&1 | grep timeline | cut -c1-170 > $O/timeline.txt
cat $O/timeline.txt
CASES="mnist_b20_cd1 dbn1000_b20_cd1 sm_b20_cd1 mnist_b10_cd1 dbn1000_b10_pcd5"
MDBN_PATH=mid python scripts/skinny_perf.py $CASES > $O/perf_mid.txt 2>&1

cat $O/perf_mid.txt $O/perf_skinny.txt | cut -c1-260
