#!/bin/sh
# Builds the CPU-emulated copy of the library (TEST INFRASTRUCTURE, see cuda_emu.h) into tests/emu/_build/.
set -e
cd "$(dirname "$0")"
mkdir -p _build
SAN=""
[ "$1" = "tsan" ] && SAN="-fsanitize=thread -g"
g++ -std=c++20 -O2 $SAN -DNSB_EMULATE -I. -shared -fPIC -pthread -x c++ ../../nspeech_b200/csrc/nspeech_b200.cu -x c++ emu_runtime.cpp \
    -o _build/libnspeech_b200_emu${1:+_$1}.so
