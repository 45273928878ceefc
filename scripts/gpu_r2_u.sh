#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python scripts/corun_probe.py 2>&1 | tail -20
