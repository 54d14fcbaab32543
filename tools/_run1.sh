set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "example_configs and 4" 2>&1 | tail -15
python tools/kernel_compare.py cnot2 16384 3,4
python tools/kernel_compare.py cnot3 2048 3,4 2
JQ_TILE_NT=3 python tools/kernel_compare.py cnot3 2048 3,4 2
