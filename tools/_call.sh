set -x
timeout -s KILL 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -s KILL 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 200 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'])"
