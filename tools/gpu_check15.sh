timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s --timeout 800 -k 4k 2>&1 | tail -6
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
