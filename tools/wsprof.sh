#!/bin/bash
# role / phase cycle profile of ba_linearize_ws_kernel (needs build/variants/wsprof.so = make WSPROF=1)
BVIO_LIB_PATH=$PWD/build/variants/${1:-wsprof}.so timeout 300 python bench.py --steps 3 --warmup 3 --stream-frames 0 --no-cpu --no-latency --pool 8 > gpurun_out/wsprof.log 2>&1
grep "ws prof" gpurun_out/wsprof.log | tail -16
