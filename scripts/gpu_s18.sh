#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider -k "demod or iir or chain or c1 or overlapped or lr4 or fm" 2>&1 | tail -2
timeout 300 python scripts/microbench.py dec chain fm lp 2>&1 | cut -c1-100
