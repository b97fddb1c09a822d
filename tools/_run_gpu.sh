set -x
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_s2.json 2> gpurun_out/bench_r1_s2.err
python tools/bench_paths.py --iters 30 > gpurun_out/paths_r1_s2.jsonl 2> gpurun_out/paths_r1_s2.err
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_s2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_s2.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_s2a.log 2>&1
python tools/bench_paths.py --iters 2 --only yolov8,yolov7,yolov3 > gpurun_out/plain_s2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'decode_tma|anchor_stream' -s 9 -c 6 -o gpurun_out/prof_r1_s2 -f python tools/bench_paths.py --iters 2 --only yolov8,yolov7,yolov3 > gpurun_out/ncu_s2b.log 2>&1
tail -3 gpurun_out/ncu_s2b.log
cat gpurun_out/bench_r1_s2.json gpurun_out/paths_r1_s2.jsonl
