set -x
(timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/r02_gputests_9.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_9.log); tail -5 gpurun_out/r02_gputests_9.log
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_wide_epi.csv python tools/sweep_bench.py 200000 500000 4,31 > gpurun_out/r02_sweep_under_ncu.log 2>&1)
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/r02_launches_wide_epi.csv') if not l.startswith('==')]
r=csv.DictReader(lines)
d=collections.defaultdict(list)
for row in r:
    n=row.get('Kernel Name','')
    if 'recomb' in n or 'split' in n or 'umma' in n or 'col_stats' in n:
        v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
        if unit in ('ns','nsecond'): v/=1e6
        elif unit in ('us','usecond'): v/=1e3
        elif unit in ('s','second'): v*=1e3
        d[n[:34]+' grid='+row['Grid Size']].append(round(v,3))
for k,v in d.items(): print(k, len(v), sorted(set(v))[:3], sorted(set(v))[-3:])
PY
