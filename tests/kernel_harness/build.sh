#!/bin/sh
# Builds the kernel-logic harness: the *same* per-lane kernel bodies (assistive_gym_b200/csrc/*.cuh)
# compiled for the host, so kernel logic can be checked against the CPU oracle on a box without a
# GPU.  Test aid only: the package never loads this library (capi.load_library loads csrc/libagphys.so).
set -e
cd "$(dirname "$0")"
/usr/bin/g++ -x c++ -std=c++17 -O2 -g -fPIC -shared -DAG_CPU_EMU -Wall -Wno-unused-function \
  -o libagphys_emu.so ../../assistive_gym_b200/csrc/agphys.cu
