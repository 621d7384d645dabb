timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/vnms_launches.csv python tools/vnms_time.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/vnms_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:14]: print(r[4][:70], r[8], r[-1], r[-2])
PY
