#!/bin/bash
timeout 300 python scripts/profile_wave.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -k "wave or pde or expm or initial_conditions or peer" 2>&1 | tail -2
