nvidia-smi -L | head -3
python -m pytest tests/test_mgpu_gpu.py -m gpu -q -x 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -3 | cut -c1-2500
