#!/bin/bash
# Family harness: the case script exported its variables; benchmarks/test_tipc/tipc.py builds the command, runs it and writes <LOG_DIR>/*_speed.json.
set -e
cd "$(dirname "$0")/../../../../../.."
exec python benchmarks/test_tipc/tipc.py --family gpt "$@"
