# sequential decoder: sanitizer pass on a small case, then the decoder tests
set -x
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_decoder.py -x -q -k "tables_match and shape0" 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -s 2>&1 | tail -25
