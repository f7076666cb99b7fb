#!/bin/bash
# usage: tests/debug_fullsize_loop.sh <lib> <n>: run the full-size slab comparison n times, print one line per run
for i in $(seq 1 $2); do PDABM_LIB=$PWD/$1 python tests/debug_fullsize.py 2>&1 | grep -E "^step|no difference" | head -1 | cut -c1-150; done
