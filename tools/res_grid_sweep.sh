#!/bin/bash
# stage times of the large-matrix eigensolver over mid-size and large orders (defaults), plus its accuracy tests
EIG_ONLY=random python tools/eig_check.py 300 552 900 1400 1800 3000 | sed 's/.*dw/dw/'
python tools/eig_check.py 129 200 333 | tail -4 | sed 's/.*dw/dw/'
python -m pytest tests/test_gpu_eigh.py -q 2>&1 | tail -2
