timeout 200 python tools/tc_bias.py 2>&1 | tail -9
timeout -k 5 600 python -m pytest tests -m gpu -q -s 2>&1 | grep "max|z\|passed\|failed\|FAILED\|flip\|float64" | head -40
timeout -k 5 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -n1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['parity'])"
