# Runs ON THE GPU BOX (gpurun): one ncu --set full capture of a bench step at 4,096 samples (plain launch sequence: one launch
# per kernel; the per-sample instruction and byte counts do not depend on the time segmentation), reduced on the box to
# gpurun_out/r02_counters.json + the raw-page CSV (the .ncu-rep itself is > 64 MiB and would block the copy back).
set -e
TAG=${1:-r02}
ncu --set full --clock-control none --import-source off -k regex:"^k_|^kw_" -c 24 -o /tmp/${TAG}_step python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary --samples 4096 --pipe-max 0 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_counters.py /tmp/${TAG}_step.ncu-rep --samples 4096 --trials 200 --T 1200 --dims 2 3 1 2 2 -o gpurun_out/${TAG}_counters.json > gpurun_out/${TAG}_counters.txt
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_raw.csv
cat gpurun_out/${TAG}_counters.txt
