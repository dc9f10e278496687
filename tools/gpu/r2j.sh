#!/bin/bash
python -m pytest tests/test_full_size.py -m gpu -x -q -k least_squares 2>&1 | tail -15 > gpurun_out/r2j_lsq.log
