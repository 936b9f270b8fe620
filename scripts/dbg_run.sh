export MDBN_SKINNY_TIMING=1
echo "== release"; python scripts/skinny_perf.py ge_b10_pcd1 2>&1 | grep timeline | tail -1
export MDBN_B200_LIB=mdbn_b200/csrc/libmdbn_b200_dbg.so
for f in 0 64 128 256 512 576 704 960; do echo "== dbg flags $f"; MDBN_SKINNY_DEBUG=$f python scripts/skinny_perf.py ge_b10_pcd1 2>&1 | grep timeline | tail -1; done
