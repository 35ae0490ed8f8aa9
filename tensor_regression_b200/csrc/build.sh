#!/bin/bash
# Build libtrb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "${BASH_SOURCE[0]}")"
make -j"${TR_JOBS:-$(nproc)}" "$@"
