set -x
O=gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:resample_rows -s 1 -c 1 -o $O/r2_rows python tools/bench_gemm_forms.py > $O/r2_rows_ncu.log 2>&1
tail -3 $O/r2_rows_ncu.log
