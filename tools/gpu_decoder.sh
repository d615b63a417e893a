# sequential decoder: tests, then phase profile of the Kodak-sized decode
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -s 2>&1 | tail -6
IC_PC_DECODE_PROF=1 timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -s -k "kodak" 2>&1 | tail -5
