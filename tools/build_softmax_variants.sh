#!/usr/bin/env bash
# Variant libraries for the softmax-exponential A/B (tools/gpu_r3d.sh): scalar FFMA / FADD, packed FFMA2 / FADD2, and
# 1 / 2 / 3 of every 8 element pairs as a polynomial on the FMA pipe instead of MUFU.EX2.
set -euo pipefail
cd "$(dirname "$0")/.."
VITED_OUT_DIR=$PWD/tools/bin/sm_scalar VITED_EXTRA_FLAGS="-DVITED_SOFTMAX_PACKED=0" bash vit-ed_b200/csrc/build.sh
VITED_OUT_DIR=$PWD/tools/bin/sm_packed VITED_EXTRA_FLAGS="-DVITED_SOFTMAX_PACKED=1 -DVITED_EXP_POLY=0" bash vit-ed_b200/csrc/build.sh
VITED_OUT_DIR=$PWD/tools/bin/sm_poly1 VITED_EXTRA_FLAGS="-DVITED_EXP_POLY=1" bash vit-ed_b200/csrc/build.sh
VITED_OUT_DIR=$PWD/tools/bin/sm_poly2 VITED_EXTRA_FLAGS="-DVITED_EXP_POLY=2" bash vit-ed_b200/csrc/build.sh
VITED_OUT_DIR=$PWD/tools/bin/sm_poly3 VITED_EXTRA_FLAGS="-DVITED_EXP_POLY=3" bash vit-ed_b200/csrc/build.sh
