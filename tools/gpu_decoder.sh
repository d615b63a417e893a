# sequential decoder: tests, then phase profile of the Kodak-sized decode
timeout -k 5 120 python -m pytest tests/test_gpu_decoder.py -x -q -s -k "not kodak" 2>&1 | tail -6
timeout -k 5 120 python -m pytest tests/test_gpu_decoder.py -x -q -s -k "kodak" 2>&1 | tail -4
IC_PC_DECODE_PROF=1 timeout -k 5 120 python -m pytest tests/test_gpu_decoder.py -x -q -s -k "kodak" 2>&1 | tail -5
