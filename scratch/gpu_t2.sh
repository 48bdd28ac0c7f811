KAI_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 3 --parity off > gpurun_out/r02_bench_l3.log 2>&1
grep -v "^\[kai\] \(relay\|publish\|sweeps\)" gpurun_out/r02_bench_l3.log | tail -c 5000 | cut -c1-1800
