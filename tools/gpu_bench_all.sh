# every bench workload / mode once (N = 1), one JSON line each -> gpurun_out/bench_all.jsonl
mkdir -p gpurun_out; : > gpurun_out/bench_all.jsonl
for w in kodak24 b64_512 cfg4 cfg1; do
  timeout -k 5 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --with-decode 2>/dev/null | tail -n1 >> gpurun_out/bench_all.jsonl
done
for m in fast fp32; do
  timeout -k 5 400 python bench.py --workload kodak24 --mode $m --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n1 >> gpurun_out/bench_all.jsonl
done
timeout 300 python tools/train_time.py --graph 2>/dev/null | tail -n1 >> gpurun_out/bench_all.jsonl
timeout 300 python tools/train_time.py --mode fp32 --cpu-batch 0 2>/dev/null | tail -n1 >> gpurun_out/bench_all.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    d=json.loads(l)
    if 'metric' in d:
        print(d['config']['workload'], d['config']['mode'], round(d['value'],1),'MPix/s', round(d['ms_per_step'],2),'ms e2e',round(d['e2e']['value'],1), 'roof',round(d['roofline']['frac'],3), 'issue', round(d['roofline']['tensor_issue_frac'],3), 'decode', d['decode'] and round(d['decode']['MPix_per_s'],1), 'parity', d['parity'] and (d['parity']['symbol_mismatches'], d['parity']['symbols'], d['parity']['max_abs_dbpp']), d['clocks'])
    else:
        print(d['workload'], d['dtype'], round(d['ms_per_step'],2), 'ms', round(d['images_per_s'],1), 'img/s', d.get('cpu_baseline'))
PY
