for v in v0 v1 v2 v0 v2; do cp onda_b200/_ab/$v.so onda_b200/libonda_b200.so; timeout 200 python bench.py --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['roofline']['kernel_ms'], d['ms_per_step'])"; done
cp onda_b200/_ab/v2.so onda_b200/libonda_b200.so
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
