#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/probe.txt
: > $S
for v in "qk32 0" "pv32ss 0" "pv32ss 1" "pv32ts 0" "pv32ts 1" "pv64ss 0" "pv64ss 1" "pv64ts 0" "pv64ts 1"; do
  timeout 60 tools/bin/umma_probe $v >> $S 2>&1; echo "  exit=$?" >> $S
done
cat $S
