set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_check.py 2>&1 | grep -E "OK|FAIL|Error|error" | head
