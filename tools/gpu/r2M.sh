#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rescale_relin_galois or double_hoisted_levels" 2>&1 | tail -15 > gpurun_out/r2M_tests.log
