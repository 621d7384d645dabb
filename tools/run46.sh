timeout 900 python -m pytest tests -m gpu -x -q -k "nms or smoke or golden or fuzz" 2>&1 | tail -4
python tools/nms_time.py
GLENET_NMS_SPATIAL=0 python tools/nms_time.py
