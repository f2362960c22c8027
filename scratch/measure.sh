python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01c_bench_reference.json 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/r01c_bench_default.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --poly-order 3 --particles 300000 --no-e2e --no-cpu-baseline > gpurun_out/r01c_bench_k3.json
python bench.py --steps 3 --warmup 3 --poly-order 4 --particles 300000 --no-e2e --no-cpu-baseline > gpurun_out/r01c_bench_k4.json
python bench.py --steps 3 --warmup 3 --ipusher 1 --no-e2e --no-cpu-baseline > gpurun_out/r01c_bench_rk4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:orbit_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_r01c_k2 python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 0 --t-step 1e-5 > gpurun_out/prof_r01c_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orbit_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_r01c_rk4 python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 0 --ipusher 1 --t-step 1e-5 > gpurun_out/prof_r01c_rk4.log 2>&1
for f in gpurun_out/r01c_bench_*.json; do cut -c1-110 $f; done
