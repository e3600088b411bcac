#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_selftest.py attention_fwd 2>&1 | tail -2 | cut -c1-2200
