# round 2, call Z8: tile-transposing planes <-> NHWC kernels of the fused trunk: tests, A/B step time, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py tests/test_gpu_train_ops.py -m gpu -q -x 2>&1 | tail -n 3
for v in 0 1; do
IC_TRAIN_TT=$v timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200
done | tee gpurun_out/r2z8_train_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z8_launches_train_step.csv python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/r2z8_ncu_train.log 2>&1
python - <<'PY'
import csv,collections
f='gpurun_out/r2z8_launches_train_step.csv'
rows=[r for r in csv.reader(open(f)) if len(r)>5]
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
data=[(r[ki], float(r[vi].replace(',',''))) for r in rows[start+1:] if r[vi].replace(',','').replace('.','').isdigit()]
n=len(data)//3
d=data[-n:]
agg=collections.OrderedDict()
for k,v in d:
    k=k[:70]
    a=agg.setdefault(k,[0,0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('launches',n,'total ms',tot/1e6)
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:16]: print('%5d %9.1f us %5.1f%% %7.1f us/l %s'%(a[0],a[1]/1e3,100*a[1]/tot,a[1]/a[0]/1e3,k))
PY
