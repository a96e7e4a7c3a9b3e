for v in e0 e1 e2 e3; do cp onda_b200/_ab/$v.so onda_b200/libonda_b200.so; timeout 200 python bench.py --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['roofline']['kernel_ms'], d['ms_per_step'])"; done
cp onda_b200/_ab/e0.so onda_b200/libonda_b200.so
