#!/bin/bash
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep -v Warn | tail -6
