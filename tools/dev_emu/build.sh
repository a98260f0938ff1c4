#!/bin/sh
# Development-only host emulation of the device headers (see sglemu.cpp).  Output: build/dev_emu/emu_player
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/dev_emu
S=softglrender_b200
g++ -std=c++17 -O2 -ffp-contract=off -mfma -Iinclude -I$S/host -I$S/harness \
  -DPLAYER_BACKEND_HEADER='"cuda_backend.h"' tools/dev_emu/sglemu.cpp $S/host/Render/CUDA/RendererCUDA.cpp \
  $S/harness/cuda_backend.cpp $S/harness/trace_player.cpp $S/harness/player_main.cpp -o build/dev_emu/emu_player
echo built build/dev_emu/emu_player
