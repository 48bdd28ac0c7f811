python -m pytest tests -m gpu -q -x 2>&1 | tail -4
KAI_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_l4.log 2>&1
grep -v "^\[kai\] \(relay\|publish\|sweeps\)" gpurun_out/r02_bench_l4.log | tail -c 5200 | cut -c1-1500
