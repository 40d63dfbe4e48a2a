ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'^k_(gbuffer|shadow|denoise|resolve|bounce)' -s 36 -c 12 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_q.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_q.csv')))
i=[k for k,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[i]
for r in rows[i+1:]:
    d=dict(zip(h,r))
    print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'])
PY
