#!/bin/bash
# round 2, GPU call 21 (1 GPU): graphdump gfa1 / gfa2 / fasta with long (whole-grid copy) and reverse-complement segment bodies
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -k "graphdump" > $O/r2c21_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c21_pytest.log
echo done
