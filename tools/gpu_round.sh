#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_vit.py 2>&1 | grep -v Warning | tail -12 | cut -c1-700
