set -x
O=gpurun_out/final6
mkdir -p $O
python scripts/ncu_small.py me && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_tiny_me_k10.csv python scripts/ncu_small.py me > $O/ncu_l2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cd_tiny -s 2 -c 1 -f -o $O/tiny_me python scripts/ncu_small.py me > $O/ncu_f2.log 2>&1
MDBN_TINY_TIMING=1 python scripts/tiny_me.py 2>&1 | grep -i timeline | tail -2 > $O/tiny_timeline.txt
python scripts/mid_timeline.py 2>&1 | grep timeline | cut -c1-170 > $O/mid_timeline.txt
tail -2 $O/ncu_f2.log; cat $O/tiny_timeline.txt
