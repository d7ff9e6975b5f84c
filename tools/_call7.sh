export PYTHONUNBUFFERED=1
timeout 300 python tools/multi_check.py 2 > gpurun_out/c7_multi2.log 2>&1; echo "multi_check rc=$?"; tail -4 gpurun_out/c7_multi2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --tile 8 > gpurun_out/c7_bench2.json 2> gpurun_out/c7_bench2.err; echo "bench2 rc=$?"; tail -c 3000 gpurun_out/c7_bench2.json; tail -5 gpurun_out/c7_bench2.err
timeout 600 python bench.py --steps 10 --warmup 3 --tile 8 --no-cpu-baseline > gpurun_out/c7_bench1.json 2> gpurun_out/c7_bench1.err; echo "bench1 rc=$?"; tail -c 3000 gpurun_out/c7_bench1.json; tail -5 gpurun_out/c7_bench1.err
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -5
