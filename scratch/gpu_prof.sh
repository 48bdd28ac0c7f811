KAI_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_launch.log 2>&1
grep -v "^\[kai\] \(relay\|scanner\|publish\|sweeps\)" gpurun_out/r02_bench_launch.log | tail -c 4500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --parity off > gpurun_out/r02_ncu_bench.log 2>&1
tail -3 gpurun_out/r02_ncu_bench.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_ncu_smoke.log 2>&1
tail -3 gpurun_out/r02_ncu_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_record -s 40 -c 4 -o gpurun_out/r02_k_record python bench.py --steps 1 --warmup 1 --parity off > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
