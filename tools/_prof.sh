set -e
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_plain.log 2>&1
tail -1 gpurun_out/b_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu1.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:"roi_align_(fwd|bwd)_cl" -s 6 -c 2 -o gpurun_out/roi_r1f -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1 || true
ls -la gpurun_out/*.ncu-rep
