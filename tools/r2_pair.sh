# frame-pair kernel: tests, A/B switches, one full ncu capture
set -x
TAG=${1:-r2b}
timeout 300 python -m pytest tests/test_gpu_pair_kernel.py -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 60 python bench.py --path pair --no-e2e --cpu-seconds 1 --steps 100 > gpurun_out/${TAG}_bench_pair.json 2> gpurun_out/${TAG}_bench_pair.err
SMB_NO_TMEM_TABLES=1 timeout 60 python bench.py --path pair --no-e2e --cpu-seconds 1 --steps 100 > gpurun_out/${TAG}_bench_pair_notmem.json 2>> gpurun_out/${TAG}_bench_pair.err
timeout 60 python tools/bench_pair_variants.py > gpurun_out/${TAG}_variants.json 2>> gpurun_out/${TAG}_bench_pair.err
timeout 120 ncu --set full --clock-control none --import-source on -k regex:stft2048p -s 3 -c 1 -o gpurun_out/${TAG}_pair python bench.py --path pair --steps 3 --warmup 3 --no-e2e --cpu-seconds 0.3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
cut -c1-250 gpurun_out/${TAG}_bench_pair.json
cut -c1-250 gpurun_out/${TAG}_bench_pair_notmem.json
cat gpurun_out/${TAG}_variants.json
