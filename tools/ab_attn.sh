#!/usr/bin/env bash
# A/B timing of attention-kernel variants on the same box: each tools/variants/lib*.so vs the in-tree library, 3 rounds
for r in 1 2 3; do
  for lib in tools/variants/lib*.so snuffy_b200/libsnuffy_b200.so; do
    echo -n "$lib: "; SNUFFY_B200_LIB=$PWD/$lib python tools/prof_attn.py
  done
done
