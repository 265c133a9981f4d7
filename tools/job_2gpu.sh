# two GPUs of one box: the NCCL / peer-memory parity tests, then bench.py exactly as the driver launches it at N = 2
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/m_tests.log
tail -3 gpurun_out/m_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/m_bench2.json 2> gpurun_out/m_bench2.err; echo "bench2 rc=$?"
