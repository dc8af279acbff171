#!/bin/bash
# 8-GPU lines: BASELINE configs[2] strong scaling (global batch 64 -> 8 per GPU), configs[3] (512^2, global 32, T=1000),
# and the round-1 style weak-scaling line (64 per GPU) for comparison
TAG=n8 bash tools/gpu/bench_ngpu.sh 8
TAG=n8_c4 bash tools/gpu/bench_ngpu.sh 8 --res 512 --diffusion-steps 1000 --batch 32 --steps 1 --warmup 1
TAG=n8_weak bash tools/gpu/bench_ngpu.sh 8 --scaling weak --batch 64 --steps 2 --warmup 2
