#!/bin/bash
python -m pytest tests/test_seal_wire.py -x -q 2>&1 | tail -25 > gpurun_out/r2k_wire.log
