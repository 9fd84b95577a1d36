for v in "" vsum2 vtwo vboth; do
  if [ -z "$v" ]; then unset MODCHOL_LIB; else export MODCHOL_LIB=$PWD/modcholesky_b200/csrc/libmodchol_b200.$v.so; fi
  echo "=== variant ${v:-base}"
  python tools/quick_bench.py --cases se99:32:1048576,gmw81:32:1048576,se90:64:200000,se99:28:1048576 --modes fast --families uniform,wishart --reps 7
done
