timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms -c 60 --csv --log-file gpurun_out/nms_launches.csv python tools/nms_time.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/nms_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:16]: print(r[4][:50], r[8], r[-1], r[-2])
PY
