#!/bin/bash
# every configuration in a process of its own (a faulting kernel poisons the context)
cd "$(dirname "$0")"
for cfg in "0 128 8 4 104" "0 128 16 2 104" "0 64 16 4 104" "1 128 1 4 104" "1 128 4 4 104" "2 128 1 4 104" "2 128 4 4 104" "2 128 8 4 104" "2 128 16 2 104" "2 64 8 4 104" "2 64 16 4 104" "2 256 4 4 104" "2 128 8 4 4099"; do
  timeout 60 ./tma_lines $cfg 2>&1 | tail -2
done
