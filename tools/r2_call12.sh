#!/bin/bash
# round 2, GPU call 12 (1 GPU): k up to 603 (5..19 words per k-mer), graphdump gfa1 / gfa2 / fasta on the GPU -- full GPU suite,
# C3 line (no regression from the templated tile staging / longer padding)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2c12_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c12_pytest.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $O/r2c12_bench_c3.json 2> $O/r2c12_bench_c3.err
echo done
