set -x
O=gpurun_out
timeout 120 python tools/bench_gemm_forms.py --knobs > $O/r2_gemm_forms_knobs.jsonl 2> $O/r2_rows_err.log
cat $O/r2_gemm_forms_knobs.jsonl; tail -5 $O/r2_rows_err.log
timeout 100 python tools/dbg_rows.py 2>&1 | grep -v "^ " | tail
timeout 240 python -m pytest tests/test_gpu_resample.py tests/test_gpu_stream.py -x -q 2>&1 | tail -15 > $O/r2_rows_tests.txt
cat $O/r2_rows_tests.txt
