#!/bin/bash
# round 2, session AG (2 GPUs): neighbour detection: one process with contexts on both GPUs is alone; two ranks are not
python - <<'PY'
import torch, goofy_b200 as gb
torch.zeros(1, device="cuda:0"); torch.zeros(1, device="cuda:1")
print("one process, contexts on both GPUs: host_neighbours =", gb.host_neighbours())
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
GOOFY_B200_HOST_RGB=1 bash tools/gpu_r02s.sh 2
