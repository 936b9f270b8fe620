# mid kernel: parity tests, phase timeline, us per step
set -x
O=gpurun_out/mid9
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "mid or auto" 2>&1 | tail -5 > $O/tests.log
cat $O/tests.log
for w in mnist_rbm_cd1_b20 mnist_rbm_pcd1_b20; do python bench.py --workload $w --no-cpu-baseline --no-extras > $O/bench_$w.json 2>/dev/null; python -c "
import json,sys
d=json.loads(open('$O/bench_$w.json').read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('single_launch',{}).get('ms_per_step'))"; done
