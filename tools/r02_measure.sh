set -x
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-frontend --no-sweep --no-sequence --no-sharded > gpurun_out/r02_plain_step.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_step.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-frontend --no-sweep --no-sequence --no-sharded > gpurun_out/r02_ncu_step.log 2>&1
timeout 300 python tools/prof_workload.py --only reg,knn,jtj > gpurun_out/r02_prof_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_prof python tools/prof_workload.py --only reg,knn,jtj > gpurun_out/r02_prof_ncu.log 2>&1
ls -la gpurun_out/r02_*
tail -c 600 gpurun_out/r02_bench_n1.err
