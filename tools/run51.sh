timeout 900 python -m pytest tests -m gpu -x -q -k "variance or soft or vnms or smoke or nms" 2>&1 | tail -3
python tools/vnms_time.py
