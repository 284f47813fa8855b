#!/bin/sh
# Run the real kernel sources under the SIMT emulator with AddressSanitizer (compute-sanitizer is
# closed on the GPU pool): catches out-of-bounds accesses of global buffers and of the per-block
# shared memory (a heap block in the emulator).  TEST TOOLING ONLY.
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
cd "$HERE/tests/emu"
g++ -std=c++17 -O1 -g -fPIC -shared -fsanitize=address -fno-omit-frame-pointer -DNWB_EMU -I. \
    -I../../needleman-wunsch_b200/csrc -o /tmp/libnwb_emu_asan.so emu_fill.cpp emu_cuda.cpp
cd "$HERE"
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
    python tools/emu_asan_run.py
