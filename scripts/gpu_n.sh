#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
GPAR_B200_LIB=$PWD/gpar_b200/libgpar_b200_prof.so timeout 120 python scripts/prof_chain.py 4096 2>&1 | tail -22
GPAR_B200_LIB=$PWD/gpar_b200/libgpar_b200_prof.so timeout 120 python scripts/prof_chain.py 8424 2>&1 | tail -22 | cut -c1-400
