#!/bin/bash
python tools/cfg34_bench.py > gpurun_out/r2h_cfg34.jsonl 2> gpurun_out/r2h_cfg34.err
