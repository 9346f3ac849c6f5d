#!/bin/sh
# Build the CPU emulation of the kernel bodies (tests only).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
mkdir -p "$here/_build"
g++ -std=c++20 -O1 -fPIC -shared -pthread -DB2_EMU -DB2_SHARD=-1 \
    -I"$root/vkfft_b200/csrc" -I"$here" \
    "$here/emu_driver.cpp" "$root/vkfft_b200/csrc/kernel_registry.cpp" "$root/vkfft_b200/csrc/planner.cpp" \
    -o "$here/_build/libb200fft_emu.so"
