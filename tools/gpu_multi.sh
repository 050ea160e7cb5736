#!/bin/bash
# multi-GPU session: the multi-device parity tests, PCIe probes (bound / unbound), bench.py both arms at N
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/lscpu_n$N.txt
timeout 900 python -m pytest tests/test_gpu_named_shapes.py -m gpu -x -q -k "multi_gpu or two_devices" > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tools/h2d_probe.py > gpurun_out/h2d_probe_n$N.txt 2> gpurun_out/h2d_probe_n$N.err; echo "probe rc=$?"; cat gpurun_out/h2d_probe_n$N.txt
timeout 600 $TR tools/h2d_probe.py --no-bind > gpurun_out/h2d_probe_nobind_n$N.txt 2>> gpurun_out/h2d_probe_n$N.err; echo "probe nobind rc=$?"; cat gpurun_out/h2d_probe_nobind_n$N.txt
bash tools/gpu_cfg.sh $N
