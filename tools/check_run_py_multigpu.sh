#!/bin/bash
# run.py under torchrun on 2 GPUs (data-parallel training, batch-sharded eval_fid) and the same eval_fid on 1 GPU:
# the PNG folders must hold the same images (draws are independent of the number of GPUs; a sample's position in the
# batch moves its fp32 GroupNorm summation order, i.e. single uint8 levels may differ).
set -u
R="${GRAFT_REPO_ROOT:-/root/repo}"
W=$(mktemp -d)
cd "$W"
C="--model diff --prior regular --dataset synthetic --a_dim 32 --batch_size 4 --epochs 1 --save_epochs 1 --diffusion_steps 4 --synthetic_size 16 --r_seed 64 --single_phase"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
export PYTHONPATH="$R"
$T "$R/run.py" $C --mode train 2>&1 | grep -E "Epoch|rror" | tail -3
$T "$R/run.py" $C --mode eval_fid --deterministic --sampling_number 6 2>&1 | grep -E "DONE|rror|Traceback" | tail -3
python "$R/run.py" $C --mode eval_fid --deterministic --sampling_number 6 --img_folder ./imgs1 2>&1 | grep -E "DONE|rror" | tail -2
python - <<'PY'
import glob
import numpy as np
from PIL import Image
a = sorted(glob.glob("imgs/*/eval-fid-fast/*.png"))
b = sorted(glob.glob("imgs1/*/eval-fid-fast/*.png"))
diff = [np.abs(np.asarray(Image.open(x)).astype(int) - np.asarray(Image.open(y)).astype(int)) for x, y in zip(a, b)]
print("files", len(a), len(b), "max level difference:", max(int(d.max()) for d in diff),
      "pixels equal: %.4f" % float(np.mean([np.mean(d == 0) for d in diff])))
PY
