set -x
timeout 300 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stft2048 -c 20 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stft2048 -s 3 -c 1 -o gpurun_out/r1h python bench.py --steps 3 --warmup 3 > gpurun_out/r1h.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ols2048 -s 1 -c 1 -o gpurun_out/ols2048 python bench_extra.py --only fir --steps 2 > gpurun_out/ols.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:istft2048 -s 1 -c 1 -o gpurun_out/istft2048 python bench_extra.py --only istft --steps 2 > gpurun_out/istft.log 2>&1
timeout 600 python bench_extra.py > gpurun_out/bench_extra.jsonl 2> gpurun_out/bench_extra_err.log
tail -1 gpurun_out/bench_line.json | cut -c1-200
