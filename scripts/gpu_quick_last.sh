timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "sliced" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-other-configs > /tmp/b.json 2> /tmp/b.err; tail -1 /tmp/b.err
python - <<'PY'
import json
d = json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g clocks %s" % (d["value"], d["e2e"]["value"], d["clocks"]))
PY
