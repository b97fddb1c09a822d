set -x
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_s3.json 2> gpurun_out/bench_r1_s3.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/plain_s3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_s3.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_s3a.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/plain_s3b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'yolov8_decode_stream|nms2_kernel' -s 6 -c 2 -o gpurun_out/prof_r1_s3 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_s3b.log 2>&1
tail -3 gpurun_out/ncu_s3b.log
cat gpurun_out/bench_r1_s3.json; cat gpurun_out/bench_r1_s3.err | tail -5
