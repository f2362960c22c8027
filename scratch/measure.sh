set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01b_bench_reference.json 2> gpurun_out/r01b_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r01b_bench_default.json 2> gpurun_out/r01b_default.err
python bench.py --steps 3 --warmup 3 --poly-order 3 --particles 300000 --no-e2e --no-cpu-baseline > gpurun_out/r01b_bench_k3.json
python bench.py --steps 3 --warmup 3 --poly-order 4 --particles 300000 --no-e2e --no-cpu-baseline > gpurun_out/r01b_bench_k4.json
python bench.py --steps 3 --warmup 3 --ipusher 1 --no-e2e --no-cpu-baseline > gpurun_out/r01b_bench_rk4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01b_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orbit_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_r01b_k3 python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 0 --poly-order 3 --particles 150000 --t-step 1e-5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:orbit_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_r01b_k4 python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 0 --poly-order 4 --particles 150000 --t-step 1e-5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:orbit_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_r01b_rk4 python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 0 --ipusher 1 --t-step 1e-5 > /dev/null 2>&1
ls -la gpurun_out | tail -12
