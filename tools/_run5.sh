for B in 1 64 296 592; do python tools/kernel_compare.py cnot2 $B 4,5,3 2; done
for B in 1 148 592; do python tools/kernel_compare.py cnot3 $B 3,5,4 1; done
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "auto_kernel or example_configs" 2>&1 | tail -4
