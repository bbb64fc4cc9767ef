#!/bin/bash
# GPU tests only (the cheapest validation)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
