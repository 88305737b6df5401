#!/bin/bash
# Kruskal first-chunk sweep: MCF 1M nodes / 10M arcs and OT 20 000^2.
for q in 1 2 4 8 16; do echo "mcf q=$q"; python bench.py --tree-only -1 --kruskal-chunk $q 2>&1 | grep -o '"kruskal": [0-9.]*'; done
for q in 8 32; do echo "ot20000 q=$q"; python bench.py --tree-only 20000 --kruskal-chunk $q 2>&1 | grep -o '"kruskal": [0-9.]*' | head -1; done
