# round 2, call Z5: issuer wait counters of every conv_tc instantiation of the inference step (IC_TC_DBG=3)
mkdir -p gpurun_out
IC_TC_DBG=3 timeout 300 python tools/hbm_kernels_once.py 24 2> gpurun_out/r2z5_dbg.txt | tail -n 1
grep "IC_TC_DBG" gpurun_out/r2z5_dbg.txt | sed 's/of \([0-9]*\) cycles/of \1 cycles/' | awk '{k=$2" "$3" "$4" "$5; c[k]++; if (c[k]<=2) print}' | head -30
