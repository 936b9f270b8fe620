set -x
O=gpurun_out/tiny7
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "tiny or aml or small" 2>&1 | tail -5 > $O/tests.log
cat $O/tests.log
timeout 300 python scripts/small_layers.py > $O/small_layers.txt 2>&1

cat $O/small_layers.txt $O/small_layers_nomreg.txt
MDBN_TINY_TIMING=1 python scripts/tiny_me.py 2>&1 | grep -i "timeline" | tail -3 | cut -c1-200
