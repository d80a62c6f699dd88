set -x
timeout 300 python scripts/c5_simplified.py 64 1 2>&1 | tail -22
