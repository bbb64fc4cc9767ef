#!/bin/bash
timeout 120 python tools/p2p_bw.py 2>&1 | tail -3
nvidia-smi topo -m 2>&1 | head -6
