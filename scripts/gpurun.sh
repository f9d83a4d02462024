#!/bin/bash
# Build in-tree first (the .so travels with the snapshot), then run a command on the B200 box.
set -e
make -C "$(dirname "$0")/../omnitokenizer_b200/csrc" -j8 | tail -1
exec /usr/local/graft/bin/gpurun "$@"
