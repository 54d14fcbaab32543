python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "5" 2>&1 | tail -5
for c in cnot2 cnot3; do python tools/kernel_compare.py $c 1 4,5,3 2; JQ_LAT_PIPE=0 python tools/kernel_compare.py $c 1 5 2; done
for c in risk_neutral cnot1 rabi; do python tools/kernel_compare.py $c 1 3,5 2; done
python tools/kernel_compare.py cnot2 148 4,5 2
python tools/kernel_compare.py cnot2 296 4,5 2
python tools/kernel_compare.py cnot3 148 3,5 1
python tools/kernel_compare.py cnot3 296 3,5,4 1
