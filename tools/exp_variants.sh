for v in "" nopf "" nopf; do
  if [ -z "$v" ]; then unset MODCHOL_LIB; else export MODCHOL_LIB=$PWD/modcholesky_b200/csrc/libmodchol_b200.$v.so; fi
  echo "=== variant ${v:-base}"
  python tools/quick_bench.py --cases se99:32:1048576,gmw81:32:1048576,se99:24:1048576 --modes fast --families uniform,wishart --reps 9
done
