bash tools/gpu_t3.sh h
python tools/prof_classes.py fp32 2>&1 | tail -8
