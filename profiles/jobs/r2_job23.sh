PPO_B=196608 PPO_ITERS=3 python profiles/prof_minibatch.py
PPO_B=196608 PPO_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_mb196608_launches_raw.csv python profiles/prof_minibatch.py > gpurun_out/r2_mb_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_mb196608_launches_raw.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
data=[(r[ki][:60], float(r[vi].replace(',',''))) for r in rows[1:]]
n=len(data)//3
for k,v in data[2*n:]: print("%-60s %9.1f us" % (k, v/1000))
print("total", sum(v for k,v in data[2*n:])/1000)
PY
