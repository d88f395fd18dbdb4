#!/bin/bash
timeout 300 python scripts/profile_wave.py 2>&1 | tail -2
