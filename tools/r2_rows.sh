set -x
O=gpurun_out
timeout 60 python tools/dbg_rows.py 2>&1 | grep -v "^ " | tail
timeout 200 python tools/bench_gemm_forms.py --knobs > $O/r2_gemm_forms_knobs.jsonl 2> $O/r2_rows_err.log
python - <<PY
import json
for l in open("$O/r2_gemm_forms_knobs.jsonl"):
    d=json.loads(l); print(d["pair"], d["clips"], d["rows_ms"], d["windows_ms"], d["rows_hbm_frac"], d["max_diff_of_peak"], d.get("rows_ms_with"))
PY
tail -5 $O/r2_rows_err.log
SMB_ROWS_SMEM_A=1 timeout 200 python tools/bench_gemm_forms.py > $O/r2_gemm_forms_smem_a.jsonl 2>> $O/r2_rows_err.log
cut -c1-30 $O/r2_gemm_forms_smem_a.jsonl | head -1; python - <<PY
import json
for l in open("$O/r2_gemm_forms_smem_a.jsonl"):
    d=json.loads(l); print("smemA", d["pair"], d["rows_ms"])
PY
timeout 400 python -m pytest tests/test_gpu_resample.py tests/test_gpu_stream.py -x -q 2>&1 | tail -15 > $O/r2_rows_tests.txt
cat $O/r2_rows_tests.txt
