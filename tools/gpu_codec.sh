python tools/codec_time.py 1 2>&1 | tail -1
python tools/codec_time.py 24 2>&1 | tail -1
