#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for p in 1 0; do echo "prefetch_desc=$p"; EB200_TMA_PREFETCH_DESC=$p python tools/probe.py wabef 2>&1 | tail -1; done
done | tee gpurun_out/r2y_prefetch_desc.txt
