#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python scripts/stream_race_check.py 2>&1 | tail -8
