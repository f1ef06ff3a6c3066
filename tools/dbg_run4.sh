echo "== NOTMA"; LSSVC_H2_NOTMA=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 2>&1 | tail -3
echo "== TMA"; timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 2>&1 | tail -8
