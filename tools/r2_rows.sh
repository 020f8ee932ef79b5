set -x
O=gpurun_out
timeout 250 python tools/measure_thdn.py 2>&1 | grep -v "windows\|smemA" | tail -6
timeout 100 python tools/check_gemm_forms.py 2>&1 | grep -v "^ " | tail -9
timeout 200 python tools/bench_gemm_forms.py --knobs > $O/r2_gemm_forms_knobs.jsonl 2> $O/r2_rows_err.log
python - <<PY
import json
for l in open("$O/r2_gemm_forms_knobs.jsonl"):
    d=json.loads(l); print(d["pair"], d["clips"], d["rows_ms"], d["windows_ms"], d["rows_hbm_frac"], d["max_diff_of_peak"], d.get("rows_ms_with"))
PY
tail -5 $O/r2_rows_err.log
timeout 600 python -m pytest tests/test_gpu_resample.py tests/test_gpu_stream.py tests/test_gpu_resample_quality.py -x -q 2>&1 | tail -4
