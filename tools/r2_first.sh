set -x
timeout 900 python -m pytest tests/test_gpu_pair_kernel.py -x -q 2>&1 | tail -15 > gpurun_out/r2a_tests.log
timeout 200 python bench.py --path fast --no-e2e --cpu-seconds 1 --steps 100 > gpurun_out/r2a_bench_fast.json 2> gpurun_out/r2a_bench_fast.err
timeout 200 python bench.py --path pair --no-e2e --cpu-seconds 1 --steps 100 > gpurun_out/r2a_bench_pair.json 2> gpurun_out/r2a_bench_pair.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stft2048p -s 3 -c 1 -o gpurun_out/r2a_pair python bench.py --path pair --steps 3 --warmup 3 --no-e2e --cpu-seconds 0.3 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_tests.log
cut -c1-400 gpurun_out/r2a_bench_fast.json
cut -c1-400 gpurun_out/r2a_bench_pair.json
