python -m pytest tests -m gpu -q -x 2>&1 | tail -6
KAI_PROFILE=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_l2.log 2>&1
grep -v "^\[kai\] \(relay\|scanner\|publish\|sweeps\)" gpurun_out/r02_bench_l2.log | tail -c 6000 | cut -c1-2600
