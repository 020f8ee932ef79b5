set -x
O=gpurun_out
timeout 200 python tools/bench_gemm_forms.py > $O/r2_gemm_forms.jsonl 2> $O/r2_rows_err.log
python - <<PY
import json
for l in open("$O/r2_gemm_forms.jsonl"):
    d=json.loads(l); print(d["pair"], d["clips"], d["rows_ms"], d["windows_ms"], d["rows_hbm_frac"], d["max_diff_of_peak"])
PY
tail -5 $O/r2_rows_err.log
timeout 400 python -m pytest tests/test_gpu_resample.py tests/test_gpu_stream.py -x -q 2>&1 | tail -15 > $O/r2_rows_tests.txt
cat $O/r2_rows_tests.txt
