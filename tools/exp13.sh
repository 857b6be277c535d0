#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/perf_rz_gta.py rz 40 64 2>&1 | tail -1
timeout 300 python tools/perf_rz_gta.py rz 80 64 2>&1 | tail -1
timeout 300 python tools/perf_rz_gta.py gta 20 16 2>&1 | tail -1
timeout 300 python tools/perf_rz_gta.py gtarz 40 16 2>&1 | tail -1
