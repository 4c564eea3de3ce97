timeout 600 python -m pytest tests -m gpu -x -q -k "mems or smoke or seed_finder_api" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from psi_b200 import capi
dev = torch.device("cuda", 0)
bench.K, bench.READ_LEN = 20, 150
W = bench.Workload(torch, dev, 0, "chr22", 20, 150, 1_000_000, 1, bench.N_PATHS, 2, 0)
print(bench.mem_bench(torch, dev, capi, W))
PY
