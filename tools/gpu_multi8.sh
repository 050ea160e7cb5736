#!/bin/bash
# 8-GPU box: multi-device parity tests, PCIe probes at 2/4/8 ranks, bench.py (both arms) at 1/2/4/8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/lscpu.txt; free -g | head -2 >> gpurun_out/lscpu.txt
timeout 900 python -m pytest tests/test_gpu_named_shapes.py -m gpu -x -q -k "multi_gpu or two_devices" > gpurun_out/pytest_multi_n8.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_n8.log
for N in 2 4 8; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N"
  timeout 600 $TR tools/h2d_probe.py > gpurun_out/h2d_probe_n$N.txt 2> gpurun_out/h2d_probe_n$N.err; echo "probe N=$N rc=$?"; cat gpurun_out/h2d_probe_n$N.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 tools/h2d_probe.py --no-bind > gpurun_out/h2d_probe_nobind_n8.txt 2>> gpurun_out/h2d_probe_n8.err; cat gpurun_out/h2d_probe_nobind_n8.txt
for N in 8 4 2 1; do bash tools/gpu_cfg.sh $N > gpurun_out/cfg_n$N.txt 2>&1; echo "bench N=$N done"; grep -E "MP/s|ERR" gpurun_out/cfg_n$N.txt | cut -c1-200; done
