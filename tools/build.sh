#!/bin/bash
# Rebuild libvecgo_cuda.so without importing the package first (a stale .so that lacks a new symbol cannot be imported).
cd "$(dirname "$0")/.." && python -c "import __graft_entry__ as g; print(g._load_builder().build(force='--force' in __import__('sys').argv))" "$@"
