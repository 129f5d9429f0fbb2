#!/bin/bash
# what does the GPU box look like?
nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm,clocks.max.mem --format=csv
lscpu | grep -E "Model name|Socket|Thread|NUMA node\(s\)"
df -h /tmp /root 2>/dev/null | head -5
ls /root/reference 2>&1 | head -2
