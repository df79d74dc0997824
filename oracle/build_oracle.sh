#!/bin/bash
# TEST INFRASTRUCTURE ONLY: compiles the plain-C restatement (oracle/oracle.c) into oracle/_build/liboracle.so.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/_build"
gcc -std=gnu99 -O2 -ffp-contract=off -fPIC -shared -o "$HERE/_build/liboracle.so" "$HERE/oracle.c" -lm
echo "built $HERE/_build/liboracle.so"
