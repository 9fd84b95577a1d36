for b in 2048 8192 16384 32768 65536 131072; do
for t in 16 0; do
  echo "=== batch $b MODCHOL_TPS_MAX=$t"
  MODCHOL_SMALL_KERNEL= MODCHOL_TPS_MAX=$t python tools/quick_bench.py --cases se99:10:$b,se99:12:$b,se99:16:$b --modes strict --families uniform --reps 9 | awk '{print $1,$2,$3,$6,$7,$8,$9,$10}'
done; done
