timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for a in "--label-block 8 --logit-margin 4" "--label-block 64 --logit-margin 30" "--label-block 16 --logit-margin 12"; do timeout 200 python bench.py --steps 40 $a 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a', d['roofline']['kernel_ms'], d['ms_per_step'])"; done
