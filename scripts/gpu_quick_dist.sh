timeout 600 python -m pytest tests -m gpu -x -q -k "distance" 2>&1 | tail -2
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from psi_b200 import capi
g = bench.build_graph("chr22")
print(bench.distance_bench(torch, torch.device("cuda", 0), capi, g, 300, 500))
PY
