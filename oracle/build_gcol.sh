#!/bin/bash
# Test infrastructure (oracle/): compiles the reference's third-party graph-colouring solver gCol HybridEA
# (klindsay28/Newton-Krylov_OOC externals/gCol/HybridEA/*.cpp, the program notebooks/IRF_coloring_dev.ipynb
# cells 19-23 hand the DIMACS file to) from the sources WHERE THEY LIE under /root/reference into
# oracle/_ref/HybridEA.  g++ on the files directly — the reference's own Makefile is not run and no source is
# copied.  Used only by tests/test_colouring.py to cross-check colouring.dimacs_lines / read_solution.
set -e
ref=${1:-/root/reference}
src=$ref/externals/gCol/HybridEA
out=$(cd "$(dirname "$0")" && pwd)/_ref
[ -d "$src" ] || { echo "no gCol sources under $ref" >&2; exit 3; }
mkdir -p "$out"
g++ -std=c++11 -O2 -w -o "$out/HybridEA" "$src"/*.cpp
echo "$out/HybridEA"
