#!/usr/bin/env bash
# The emulated kernels (tests/hostdev/kernels_emul.cpp: csrc/device/*.cuh compiled for the host) under
# AddressSanitizer + UBSan: out-of-bounds indexing and undefined behaviour in the DEVICE code show up here, on the CPU,
# before they would corrupt memory on a GPU.  Run from the repo root; restores the normal test library afterwards.
set -eu
B=tests/hostdev/_build
mkdir -p "$B" /tmp/ecm_asan
python -c "from tests.test_hostdev_kernels import load_emu; load_emu()"   # make sure the normal library exists
cp "$B/libkernels_emul.so" /tmp/ecm_asan/orig.so
trap 'cp /tmp/ecm_asan/orig.so "$B/libkernels_emul.so"' EXIT
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -fno-gnu-unique -fPIC -shared -I tests/hostdev/shim \
    -o "$B/libkernels_emul.so" tests/hostdev/kernels_emul.cpp
LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
    python -m pytest tests/test_hostdev_kernels.py tests/test_hostdev_kdtree.py tests/test_hostdev_planner.py tests/test_hostdev_longpaths.py \
    tests/test_window_locality.py -x -q -k "not gloo"
