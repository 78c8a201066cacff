#!/bin/bash
# build both libraries, then run a command on the B200 box:  tools/gpu.sh [--timeout S] [--gpus N] -- 'command'
set -e
cd "$(dirname "$0")/.."
make -s -j8 -C lvt_b200/csrc > /dev/null
make -s -C oracle > /dev/null
exec /usr/local/graft/bin/gpurun "$@"
