# runs the single-layer tcgen05 checks, each under its own timeout so a hang cannot eat the box
for cfg in "1 16 8 fast" "1 16 8 exact" "1 16 16 exact" "1 32 32 exact res" "2 40 40 exact res" "3 48 24 fast res" "1 192 128 exact res"; do
  echo "== $cfg"
  timeout -k 5 90 python tools/tc_debug.py $cfg 2>&1 | tail -12
  echo "rc=$?"
done
