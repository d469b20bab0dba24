#!/bin/bash
# Rebuilds everything that travels to the GPU box (product library, host-stepped test build, data generator), then runs the
# given gpurun arguments.  usage: tools/gpu.sh --timeout 1200 -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -m pansvr_b200.build > /dev/null
make -s -C tests/emul "$(pwd)/tests/emul/fc_aln_emul" 2>&1 | grep -E "error" || true
make -s -C tests/emul
make -s -C oracle all > /dev/null
python -c "from benchdata import config3; config3.build_generator()"
exec /usr/local/graft/bin/gpurun "$@"
