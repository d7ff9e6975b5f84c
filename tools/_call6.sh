export PYTHONUNBUFFERED=1
nvidia-smi -L
timeout 300 python tools/multi_check.py 2 > gpurun_out/c6_multi2.log 2>&1; echo "multi_check rc=$?"; tail -12 gpurun_out/c6_multi2.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py > gpurun_out/c6_dist2.log 2>&1; echo "dist_check rc=$?"; tail -12 gpurun_out/c6_dist2.log | cut -c1-400
timeout 300 python tools/multi_check.py 2 ecdsa+secp256k1 > gpurun_out/c6_multi2_ecdsa.log 2>&1; echo "ecdsa rc=$?"; tail -3 gpurun_out/c6_multi2_ecdsa.log | cut -c1-600
