export MDBN_SKINNY_TIMING=1
export MDBN_B200_LIB=mdbn_b200/csrc/libmdbn_b200_dbg.so
for f in ${FLAGS:-0 4 8 16 20 28}; do echo "== dbg flags $f"; MDBN_SKINNY_DEBUG=$f python scripts/skinny_perf.py ${CASE:-ge_b10_pcd1} 2>&1 | grep timeline | tail -1; done
